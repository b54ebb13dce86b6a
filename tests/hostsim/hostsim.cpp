// TEST INFRASTRUCTURE ONLY.  Compiles the *device* per-voxel code (svb_models_asl_b200/csrc/voxel_step.h,
// model_*.h, philox.h) with the host compiler so that the kernel arithmetic can be checked against the
// oracle in the GPU-less build container (pytest -m "not gpu").  Never linked into libsvbasl.so and never
// imported by the product package; on the GPU box the parity tests go through the real C ABI instead.
#include <cstdio>
#include <cstring>

#include "voxel_step.h"
#include "model_aslrest.h"
#include <vector>
#include "model_nn.h"
#include "model_disp.h"
#include "model_list.h"

using namespace svb;

// FL: flavour of VoxelStep (0 generic, 1 lean, 2 lean + spatial prior).  Flavour 2 also stages the neighbours'
// samples of the first spatial parameter in an NbTile, as the CUDA kernel does in shared memory.
template <class M, int NBT, int FL = 0>
static int run_step(const svbasl_model *md, const svbasl_engine *e, const svbasl_adam *ad, int64_t step0, float *cost,
                    float *grad, double *cost_sum, double *ak_grad) {
    const int n_iters = ad ? ad->n_iters : 1;
    const DevModel dm = make_dev_model(*md);
    const EngineConst ec = make_engine_const(*e);
    for (int64_t local = 0; local < e->n_vox; ++local) {
        const int64_t w = e->w_begin + local;
        VoxelStep<M, NBT, FL> vs;
        vs.load(*e, w);
        NbTile nbt = {nullptr, 0, -1};
        std::vector<float> tile;
        if (FL == 2) {
            for (int i = 0; i < e->n_par && nbt.param < 0; ++i)
                if (e->prior_type[i] == SVBASL_PRIOR_MRF) nbt.param = i;
            if (nbt.param >= 0) {
                const int S = e->n_samples;
                tile.assign((size_t)6 * S, 0.0f);
                const float *src = e->spatial_samples + (int64_t)ec.sp_slot[nbt.param] * S * e->ld;
                for (int k = 0; k < 6; ++k) {
                    int u = e->neighbours[(int64_t)k * e->ld + w];
                    if (u < 0) u = (int)w;                       // a missing neighbour is the voxel itself: zero difference
                    for (int sidx = 0; sidx < S; ++sidx) tile[(size_t)sidx * 6 + k] = src[(int64_t)sidx * e->ld + u];
                }
                nbt.v = tile.data();
                nbt.stride = 1;
            }
        }
        for (int it = 0; it < n_iters; ++it) {
            const int64_t step = (ad ? ad->step0 : step0) + it;
            const int row0 = (ad && ad->n_batches > 1) ? (int)(step % ad->n_batches) : e->t_row0;
            float c = vs.elbo_grad(dm, *e, ec, w, step, row0, nbt);
            if (cost) cost[w] = c;
            if (grad) vs.store_grads(*e, grad, w);
            if (ad) {
                if (vs.grads_finite() && c == c) vs.adam_update(*e, *ad, ad->lr_t[step], w, it == n_iters - 1, ad->m + w, ad->v + w, e->ld, ad->m + w, ad->v + w, e->ld);
                else { if (it == n_iters - 1) vs.store_state(*e, w); c = 0.0f; }
                if (FL != 1 && it == n_iters - 1 && e->spatial_samples_out) vs.store_next_samples(*e, ec, w, step + 1);
            }
            if (cost_sum) cost_sum[it] += c;
            if (ak_grad)
                for (int i = 0; i < VoxelStep<M, NBT>::N; ++i)
                    if (e->prior_type[i] == SVBASL_PRIOR_MRF) ak_grad[ec.sp_slot[i]] += vs.ak_out[i];
        }
    }
    return 0;
}

template <class M>
static int run_eval(const svbasl_model *md, const float *params, const float *tpts, float *out, int64_t n_rows,
                    int n_samples, int n_batch, int64_t n_t_rows) {
    const int64_t rows_per_t = n_rows / n_t_rows;
    const DevModel dm = make_dev_model(*md);
    for (int64_t row = 0; row < n_rows; ++row) {
        float x[M::P > 0 ? M::P : 1];
        for (int p = 0; p < M::P; ++p) x[p] = params[(int64_t)p * n_rows + row];
        typename M::Vox vx = M::load_vox(dm, row / n_samples);
        for (int b = 0; b < n_batch; ++b)
            out[row * n_batch + b] = M::predict(dm, vx, x, tpts[(row / rows_per_t) * n_batch + b]);
    }
    return 0;
}

static uint32_t canon(uint32_t f) {
    if (f & SVBASL_F_ARTONLY) f |= SVBASL_F_INFERART;
    if (f & SVBASL_F_ARTONLY) f &= ~(uint32_t)(SVBASL_F_INCWM | SVBASL_F_INFERWM);
    return f;
}

static bool is_nn(const svbasl_model *md) { return md->kind == SVBASL_MODEL_ASLNN; }
static uint32_t canon_disp(uint32_t f) {
    f &= (SVBASL_F_CASL | SVBASL_F_INFERATT | SVBASL_F_INFERART | SVBASL_F_ARTONLY | SVBASL_F_DISP_INFER);
    if (f & SVBASL_F_ARTONLY) f |= SVBASL_F_INFERART;
    return f;
}

extern "C" int hostsim_n_params(const svbasl_model *md) {
    if (is_nn(md)) return AslNN::P;
    if (md->kind == SVBASL_MODEL_ASLREST_DISP) {
        const uint32_t fd = canon_disp(md->flags);
#define X(F) if (fd == F) return AslDisp<F>::P;
        HOSTSIM_DISP_FLAGS
#undef X
        return -1;
    }
    const uint32_t f = canon(md->flags);
#define X(F) if (md->kind == SVBASL_MODEL_ASLREST && f == F) return AslRest<F>::P;
    HOSTSIM_ASLREST_FLAGS
#undef X
    return -1;
}

extern "C" int hostsim_evaluate(const svbasl_model *md, const float *params, const float *tpts, float *out,
                                int64_t n_rows, int n_samples, int n_batch, int64_t n_t_rows) {
    if (is_nn(md)) return run_eval<AslNN>(md, params, tpts, out, n_rows, n_samples, n_batch, n_t_rows);
    if (md->kind == SVBASL_MODEL_ASLREST_DISP) {
        const uint32_t fd = canon_disp(md->flags);
#define X(F) if (fd == F) return run_eval<AslDisp<F>>(md, params, tpts, out, n_rows, n_samples, n_batch, n_t_rows);
        HOSTSIM_DISP_FLAGS
#undef X
        return -2;
    }
    const uint32_t f = canon(md->flags);
#define X(F) if (md->kind == SVBASL_MODEL_ASLREST && f == F) return run_eval<AslRest<F>>(md, params, tpts, out, n_rows, n_samples, n_batch, n_t_rows);
    HOSTSIM_ASLREST_FLAGS
#undef X
    return -2;
}

// nbt: 0 = dynamic batch loop, 6 = register-resident batch (only for the fast-path layouts)
extern "C" int hostsim_step(const svbasl_model *md, const svbasl_engine *e, const svbasl_adam *ad, int64_t step,
                            float *cost, float *grad, double *cost_sum, double *ak_grad, int nbt) {
    if (is_nn(md)) {
        if (nbt == 6) return run_step<AslNN, 6>(md, e, ad, step, cost, grad, cost_sum, ak_grad);
        if (nbt == 106) return run_step<AslNN, 6, 1>(md, e, ad, step, cost, grad, cost_sum, ak_grad);
        return run_step<AslNN, 0>(md, e, ad, step, cost, grad, cost_sum, ak_grad);
    }
    if (md->kind == SVBASL_MODEL_ASLREST_DISP) {
        const uint32_t fd = canon_disp(md->flags);
#define X(F) if (fd == F) return run_step<AslDisp<F>, 0>(md, e, ad, step, cost, grad, cost_sum, ak_grad);
        HOSTSIM_DISP_FLAGS
#undef X
        return -2;
    }
    const uint32_t f = canon(md->flags);
    if (nbt == 0) {
#define X(F) if (md->kind == SVBASL_MODEL_ASLREST && f == F) return run_step<AslRest<F>, 0>(md, e, ad, step, cost, grad, cost_sum, ak_grad);
        HOSTSIM_ASLREST_FLAGS
#undef X
    }
    // nbt = 100 + NBT selects the lean (production) flavour of the same layout, 200 + NBT lean + spatial
#define Y(F, NBT) if (md->kind == SVBASL_MODEL_ASLREST && f == F && nbt == NBT) return run_step<AslRest<F>, NBT>(md, e, ad, step, cost, grad, cost_sum, ak_grad); \
    if (md->kind == SVBASL_MODEL_ASLREST && f == F && nbt == 100 + NBT) return run_step<AslRest<F>, NBT, 1>(md, e, ad, step, cost, grad, cost_sum, ak_grad); \
    if (md->kind == SVBASL_MODEL_ASLREST && f == F && nbt == 200 + NBT) return run_step<AslRest<F>, NBT, 2>(md, e, ad, step, cost, grad, cost_sum, ak_grad);
    HOSTSIM_FAST
#undef Y
    return -2;
}

extern "C" void hostsim_sample_spatial(const svbasl_engine *e, int64_t n_local, int64_t step, float *out) {
    const EngineConst ec = make_engine_const(*e);
    const uint32_t key = rng_key(e->seed, step);
    for (int64_t u = 0; u < n_local; ++u)
        for (int p = 0; p < e->n_par; ++p) {
            if (ec.sp_slot[p] < 0) continue;
            for (int s = 0; s < e->n_samples; ++s)
                out[((int64_t)ec.sp_slot[p] * e->n_samples + s) * e->ld + u] = sample_theta(*e, key, u, p, s);
        }
}

extern "C" void hostsim_fill_eps(float *eps, int64_t n_vox, int64_t ld, int64_t vox_offset, int n_par, int n_samples,
                                 uint64_t seed, int64_t step) {
    const uint32_t key = rng_key(seed, step);
    for (int64_t w = 0; w < n_vox; ++w)
        for (int j = 0; j < n_par; ++j)
            for (int s = 0; s < n_samples; ++s)
                eps[((int64_t)j * n_samples + s) * ld + w] = normal_at(key, vox_offset + w, j, s, n_samples);
}

// Q(a, x_k), dQ/da, dQ/dx along a sequence of arguments through the running evaluator (model_disp.h:
// gamma_run_eval) and, for comparison, through the from-scratch igammac_d at every point.
extern "C" void hostsim_gamma_run(float a, const float *xs, int n, float *q, float *dqa, float *dqx, float *q_ref,
                                  float *dqa_ref, float *dqx_ref) {
    const GammaConst g = gamma_const(a);
    GammaRun r = gamma_run_start();
    for (int k = 0; k < n; ++k) {
        const float lnx = logf(fmaxf(xs[k], 1e-30f));
        gamma_run_eval(g, r, xs[k], lnx, q[k], dqa[k], dqx[k]);
        igammac_d(g, xs[k], lnx, q_ref[k], dqa_ref[k], dqx_ref[k]);
    }
}
