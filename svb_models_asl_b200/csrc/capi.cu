// capi.cu - the extern "C" surface declared in include/svbasl.h: argument validation, dispatch on the
// model layout to the matching kernel instantiation, and the small stand-alone kernels (RNG fill,
// initialisation statistics, hyper-parameter Adam, host-staged iteration).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "kernels.cuh"
#include "_gen/groups.h"

namespace svb {

// csrc/nn_tc.cu
struct NnTcArgs;
void nn_tc_build_b_tile(const NNWeights &w, float *tile);
int nn_tc_launch(const DevModel &dm, const float *b_tile, const float *params, const float *tpts, float *out,
                 float *hidden, int64_t n_rows, int32_t n_batch, int64_t n_t_rows, int32_t *status, cudaStream_t st);

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static uint32_t canonical_flags(const svbasl_model *m) {
    uint32_t f = m->flags;
    if (m->kind == SVBASL_MODEL_ASLREST) {
        f &= (SVBASL_F_CASL | SVBASL_F_INFERATT | SVBASL_F_INFERART | SVBASL_F_INCWM | SVBASL_F_INFERWM |
              SVBASL_F_INFERT1 | SVBASL_F_ARTONLY);
        if (f & SVBASL_F_ARTONLY) f |= SVBASL_F_INFERART;                      // aslrest.py:137-138
        // inferwm does NOT imply incwm (only pvcorr sets both, aslrest.py:103-105): the layout keeps them apart
        if (f & SVBASL_F_ARTONLY) f &= ~(uint32_t)(SVBASL_F_INCWM | SVBASL_F_INFERWM);
    } else if (m->kind == SVBASL_MODEL_ASLREST_DISP) {
        f &= (SVBASL_F_CASL | SVBASL_F_INFERATT | SVBASL_F_INFERART | SVBASL_F_ARTONLY | SVBASL_F_DISP_INFER);
        if (f & SVBASL_F_ARTONLY) f |= SVBASL_F_INFERART;
    } else if (m->kind == SVBASL_MODEL_ASLNN) {
        f &= SVBASL_F_NN_TC;
    }
    return f;
}

// Best kernel for (model layout, batch size): prefers the compile-time batch size, and the lean (production)
// flavour when the call allows it.
// `flavour`: 0 = generic only, 1 = lean allowed (no spatial prior in use), 2 = lean-spatial allowed.
static const KernelEntry *find_entry(const svbasl_model *m, int nbt, bool want_eval, int flavour = 0, bool allow_tc = true) {
    uint32_t f = canonical_flags(m);
    if (m->kind == SVBASL_MODEL_ASLNN && (want_eval || !allow_tc)) f &= ~(uint32_t)SVBASL_F_NN_TC;
    const KernelEntry *best = nullptr;
    int best_score = -1;
    for (int g = 0; g < kNumEntryGroups; ++g) {
        for (const KernelEntry *e = kEntryGroups[g](); e->kind >= 0; ++e) {
            if (e->kind != m->kind || e->flags != f) continue;
            if (want_eval) {
                if (e->eval) return e;
                continue;
            }
            if (e->nbt != nbt && e->nbt != 0) continue;
            if (e->lean && e->lean != flavour) continue;
            const int score = (e->nbt == nbt ? 2 : 0) + (e->lean ? 1 : 0);
            if (score > best_score) { best = e; best_score = score; }
        }
    }
    // no tensor-core instantiation for this call: the FP32-pipe kernel computes the same thing
    if (!best && allow_tc && (f & SVBASL_F_NN_TC) && m->kind == SVBASL_MODEL_ASLNN) return find_entry(m, nbt, want_eval, flavour, false);
    return best;
}

static uint32_t mrf_mask(const svbasl_engine *e) {
    uint32_t mask = 0;
    for (int i = 0; i < e->n_par && i < SVBASL_MAX_PAR; ++i)
        if (e->prior_type[i] == SVBASL_PRIOR_MRF) mask |= 1u << i;
    return mask;
}

static int validate(const svbasl_model *m, const svbasl_engine *e, const KernelEntry **entry, bool lean_ok) {
    if (!m || !e) { set_error("null descriptor"); return SVBASL_E_INVALID; }
    if (e->n_vox < 0 || e->w_begin < 0 || e->ld < e->w_begin + e->n_vox) {
        set_error("bad extents: n_vox=%lld w_begin=%lld ld=%lld", (long long)e->n_vox, (long long)e->w_begin, (long long)e->ld);
        return SVBASL_E_INVALID;
    }
    if (e->n_samples < 1 || e->n_batch < 1 || e->t_full < 1 || e->t_row_stride < 1) {
        set_error("bad sizes: S=%d B=%d T=%d row_stride=%d", e->n_samples, e->n_batch, e->t_full, e->t_row_stride);
        return SVBASL_E_INVALID;
    }
    if (!e->state || !e->data || (!e->tpts && !e->ti)) { set_error("state, data and tpts|ti are required"); return SVBASL_E_INVALID; }
    const uint32_t mask = mrf_mask(e);
    if (mask && e->latent != SVBASL_LATENT_NUMERIC) {
        set_error("spatial (M) priors need the sample-based latent loss");
        return SVBASL_E_INVALID;
    }
    if (mask && (!e->neighbours || !e->log_ak || !e->spatial_samples)) {
        set_error("spatial prior without neighbours / log_ak / spatial_samples (call svbasl_sample_spatial first)");
        return SVBASL_E_INVALID;
    }
    const bool production = lean_ok && !e->eps && e->latent == SVBASL_LATENT_NUMERIC;
    const KernelEntry *k = find_entry(m, e->n_batch, false, production ? (mask ? 2 : 1) : 0);
    if (!k) {
        set_error("no kernel compiled for model kind=%d flags=0x%x", m->kind, canonical_flags(m));
        return SVBASL_E_UNSUPPORTED;
    }
    if (e->n_par != k->n_params + 1) {
        set_error("engine n_par=%d but the model has %d parameters (+1 noise)", e->n_par, k->n_params);
        return SVBASL_E_INVALID;
    }
    if (m->kind == SVBASL_MODEL_ASLREST_DISP) {
        if (m->flags & (SVBASL_F_INCWM | SVBASL_F_INFERWM | SVBASL_F_INFERT1)) {
            set_error("aslrest_disp: WM / T1 inference is not supported with dispersion");
            return SVBASL_E_UNSUPPORTED;
        }
        if (m->conv_nt < 2 || !(m->conv_tmax > 0.0f) || !(m->conv_dt > 0.0f)) { set_error("aslrest_disp: bad convolution grid"); return SVBASL_E_INVALID; }
    }
    if (m->kind == SVBASL_MODEL_ASLNN && !m->nn_weights) { set_error("aslnn needs nn_weights"); return SVBASL_E_INVALID; }
    if (m->kind == SVBASL_MODEL_ASLREST && (canonical_flags(m) & SVBASL_F_INFERT1) == 0 && !(m->t1 > 0.0f)) {
        set_error("t1 must be positive");
        return SVBASL_E_INVALID;
    }
    *entry = k;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
__global__ void fill_eps_kernel(float *eps, int64_t n_vox, int64_t ld, int64_t vox_offset, int n_par, int n_samples,
                                uint64_t seed, int64_t step) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_vox) return;
    const uint32_t key = rng_key(seed, step);
    // row j of samples s, s+1 = one Philox call (philox.h: stream_pair)
    for (int j = 0; j < n_par; ++j) {
        for (int s = 0; s < n_samples; s += 2) {
            float n0, n1;
            normal_pair(key, vox_offset + w, stream_pair(j, s, n_samples), n0, n1);
            eps[((int64_t)j * n_samples + s) * ld + w] = n0;
            if (s + 1 < n_samples) eps[((int64_t)j * n_samples + s + 1) * ld + w] = n1;
        }
    }
}

__global__ void init_stats_kernel(const float *data, const float *tpts, int64_t n_vox, int64_t ld, int t_full,
                                  float *mean_t, float *max_t, float *var_t, float *t_at_max) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_vox) return;
    float sum = 0.0f, mx = -INFINITY, tmx = 0.0f;
    for (int r = 0; r < t_full; ++r) {
        const float y = data[(int64_t)r * ld + w];
        sum += y;
        if (y > mx) { mx = y; tmx = tpts ? tpts[(int64_t)r * ld + w] : 0.0f; }   // first maximum (tf.argmax)
    }
    const float mean = sum / (float)t_full;
    float ss = 0.0f;
    for (int r = 0; r < t_full; ++r) {
        const float d = data[(int64_t)r * ld + w] - mean;
        ss += d * d;
    }
    if (mean_t) mean_t[w] = mean;
    if (max_t) max_t[w] = mx;
    if (var_t) var_t[w] = ss / (float)t_full;                  // tf.nn.moments: population variance
    if (t_at_max) t_at_max[w] = tmx;
}

// graph-friendly tail: lr_t from the device table at *step_dev, then zero ak_grad and advance the counter
__global__ void hyper_step_dev_kernel(float *log_ak, float *m, float *v, double *ak_grad, int n, float grad_scale,
                                      const float *lr_t, long long *step_dev, float b1, float b2, float eps) {
    const int k = threadIdx.x;
    const long long step = *step_dev;
    if (k < n) {
        const float g = (float)(ak_grad[k] * (double)grad_scale);
        const float mm = b1 * m[k] + (1.0f - b1) * g;
        const float vv = b2 * v[k] + (1.0f - b2) * g * g;
        m[k] = mm;
        v[k] = vv;
        log_ak[k] -= lr_t[step] * mm / (sqrtf(vv) + eps);
        ak_grad[k] = 0.0;
    }
    __syncthreads();
    if (k == 0) *step_dev = step + 1;
}

// svbasl_hyper_step_peers as a launch of its own (halo_mode "peer" without the fused tail): one CTA, hyper_tail
__global__ void hyper_step_peers_kernel(svbasl_hyper hy, double *ak_grad, float grad_scale) {
    __shared__ double part[SVBASL_MAX_PEERS][SVBASL_MAX_SPATIAL];
    hyper_tail(hy, ak_grad, grad_scale, part);
}

__global__ void advance_step_kernel(long long *step_dev, long long inc) { *step_dev += inc; }

__global__ void hyper_step_kernel(float *log_ak, float *m, float *v, const double *ak_grad, int n, float grad_scale,
                                  float lr_t, float b1, float b2, float eps) {
    const int k = threadIdx.x;
    if (k >= n) return;
    const float g = (float)(ak_grad[k] * (double)grad_scale);
    const float mm = b1 * m[k] + (1.0f - b1) * g;
    const float vv = b2 * v[k] + (1.0f - b2) * g * g;
    m[k] = mm;
    v[k] = vv;
    log_ak[k] -= lr_t * mm / (sqrtf(vv) + eps);
}

}  // namespace svb

using namespace svb;

struct svbasl_host_ctx {
    cudaStream_t copy_stream, run_stream;
    cudaEvent_t copied[2], consumed[2], done;
    float *d_data[2], *d_tpts[2], *d_ti[2];
    double *d_cost;
    int64_t ld;
    int32_t n_batch;
    int slot;
    long long calls;
};

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            set_error("%s: %s", #expr, cudaGetErrorString(_e));                             \
            return SVBASL_E_CUDA;                                                           \
        }                                                                                   \
    } while (0)

extern "C" {

const char *svbasl_last_error(void) { return g_err; }

int svbasl_abi_version(void) { return SVBASL_ABI_VERSION; }

int svbasl_model_n_params(const svbasl_model *model) {
    if (!model) { set_error("null model"); return SVBASL_E_INVALID; }
    const KernelEntry *k = find_entry(model, 0, true);
    if (!k) {
        set_error("no kernel compiled for model kind=%d flags=0x%x", model->kind, canonical_flags(model));
        return SVBASL_E_UNSUPPORTED;
    }
    return k->n_params;
}

int svbasl_n_state(const svbasl_model *model, const svbasl_engine *engine) {
    int p = svbasl_model_n_params(model);
    if (p < 0) return p;
    if (!engine) { set_error("null engine"); return SVBASL_E_INVALID; }
    const int n = p + 1;
    int n_ard = 0;
    for (int i = 0; i < n; ++i) n_ard += (engine->prior_type[i] == SVBASL_PRIOR_ARD);
    return 2 * n + n * (n - 1) / 2 + n_ard;
}

int svbasl_evaluate(const svbasl_model *model, const float *params, const float *tpts, float *out, int64_t n_rows,
                    int32_t n_samples, int32_t n_batch, int64_t n_t_rows, void *stream) {
    if (!model || !tpts || !out) { set_error("null argument"); return SVBASL_E_INVALID; }
    const KernelEntry *k = find_entry(model, 0, true);
    if (!k) {
        set_error("no kernel compiled for model kind=%d flags=0x%x", model->kind, canonical_flags(model));
        return SVBASL_E_UNSUPPORTED;
    }
    if (k->n_params > 0 && !params) { set_error("null params"); return SVBASL_E_INVALID; }
    if (n_rows < 0 || n_samples < 1 || n_batch < 1 || n_t_rows < 1 || (n_rows % n_t_rows) != 0 || (n_rows % n_samples) != 0) {
        set_error("bad shapes: rows=%lld S=%d B=%d t_rows=%lld", (long long)n_rows, n_samples, n_batch, (long long)n_t_rows);
        return SVBASL_E_INVALID;
    }
    EvalArgs a;
    a.md = make_dev_model(*model);
    a.params = params;
    a.tpts = tpts;
    a.out = out;
    a.n_rows = n_rows;
    a.n_t_rows = n_t_rows;
    a.n_samples = n_samples;
    a.n_batch = n_batch;
    return k->eval(a, (cudaStream_t)stream);
}

int svbasl_nn_pack_weights(const svbasl_model *model, float *host_tile) {
    if (!model || !host_tile || model->kind != SVBASL_MODEL_ASLNN || !model->nn_weights) {
        set_error("nn_pack_weights needs an aslnn model with weights");
        return SVBASL_E_INVALID;
    }
    const DevModel dm = make_dev_model(*model);
    nn_tc_build_b_tile(dm.nn, host_tile);
    return 0;
}

int svbasl_nn_evaluate_tc(const svbasl_model *model, const float *b_tile, const float *params, const float *tpts,
                          float *out, float *hidden, int64_t n_rows, int32_t n_batch, int64_t n_t_rows,
                          int32_t *status, void *stream) {
    if (!model || model->kind != SVBASL_MODEL_ASLNN || !model->nn_weights || !b_tile || !params || !tpts ||
        (!out && !hidden)) {
        set_error("bad nn_evaluate_tc arguments");
        return SVBASL_E_INVALID;
    }
    if (n_rows < 0 || n_batch < 1 || n_t_rows < 1 || (n_rows % n_t_rows) != 0) {
        set_error("bad shapes: rows=%lld B=%d t_rows=%lld", (long long)n_rows, n_batch, (long long)n_t_rows);
        return SVBASL_E_INVALID;
    }
    return nn_tc_launch(make_dev_model(*model), b_tile, params, tpts, out, hidden, n_rows, n_batch, n_t_rows, status,
                        (cudaStream_t)stream);
}

static int run_step(const svbasl_model *model, const svbasl_engine *engine, const svbasl_adam *adam, int64_t step,
                    float *cost, float *grad, double *cost_sum, long long *nan_count, void *stream,
                    const svbasl_hyper *hyper = nullptr) {
    const KernelEntry *k = nullptr;
    int rc = validate(model, engine, &k, adam != nullptr && !cost && !grad);
    if (rc) return rc;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.md = make_dev_model(*model);
    a.e = *engine;
    a.ec = make_engine_const(*engine);
    a.update = adam ? 1 : 0;
    {
        const int n = engine->n_par;
        int n_ard = 0;
        for (int i = 0; i < n; ++i) n_ard += (engine->prior_type[i] == SVBASL_PRIOR_ARD);
        a.n_state = 2 * n + n * (n - 1) / 2 + n_ard;
    }
    if (adam) {
        if (!adam->m || !adam->v || !adam->lr_t || adam->n_iters < 1 || adam->n_batches < 1) {
            set_error("bad adam descriptor");
            return SVBASL_E_INVALID;
        }
        if (mrf_mask(engine) && adam->n_iters != 1) {
            set_error("spatial priors couple neighbouring voxels: n_iters must be 1");
            return SVBASL_E_INVALID;
        }
        a.ad = *adam;
        step = adam->step0;
    }
    if (engine->spatial_samples_out) {
        if (!adam || !mrf_mask(engine) || engine->eps || engine->spatial_samples_out == engine->spatial_samples) {
            set_error("spatial_samples_out needs a fused update with a spatial prior, in-kernel draws (eps == NULL) and a "
                      "buffer different from spatial_samples");
            return SVBASL_E_INVALID;
        }
    }
    if (engine->peer_lo || engine->peer_hi) {
        const int64_t nlo = engine->peer_lo ? engine->peer_lo_count : 0, nhi = engine->peer_hi ? engine->peer_hi_count : 0;
        if (nlo < 0 || nhi < 0 || nlo + nhi > engine->n_vox ||
            (engine->peer_lo && engine->peer_lo_first != engine->w_begin) ||
            (engine->peer_hi && engine->peer_hi_first != engine->w_begin + engine->n_vox - nhi)) {
            set_error("peer_lo / peer_hi must name the first / last owned voxels of the launch");
            return SVBASL_E_INVALID;
        }
    }
    if (hyper) {
        if (!adam || !mrf_mask(engine) || !engine->ak_grad || !engine->step_dev || hyper->step_dev != engine->step_dev ||
            hyper->log_ak != engine->log_ak || !hyper->m || !hyper->v || !hyper->lr_t || !hyper->done_ctas ||
            hyper->n_spatial < 1 || hyper->n_spatial > SVBASL_MAX_SPATIAL || hyper->world < 1 ||
            hyper->world > SVBASL_MAX_PEERS || hyper->rank < 0 || hyper->rank >= hyper->world ||
            (hyper->world > 1 && !hyper->status)) {
            set_error("bad svbasl_hyper descriptor (needs a spatial prior, ak_grad, the engine's step_dev and log_ak)");
            return SVBASL_E_INVALID;
        }
        for (int r = 0; r < hyper->world && hyper->world > 1; ++r)
            if (!hyper->mailboxes[r]) { set_error("mailbox of rank %d is NULL", r); return SVBASL_E_INVALID; }
        a.hy = *hyper;
    }
    a.step = step;
    a.cost = cost;
    a.grad = grad;
    a.cost_sum = cost_sum;
    a.nan_count = nan_count;
    return k->step(a, (cudaStream_t)stream);
}

int svbasl_elbo_grad(const svbasl_model *model, const svbasl_engine *engine, int64_t step, float *cost, float *grad,
                     double *cost_sum, void *stream) {
    return run_step(model, engine, nullptr, step, cost, grad, cost_sum, nullptr, stream);
}

int svbasl_step(const svbasl_model *model, const svbasl_engine *engine, const svbasl_adam *adam, double *cost_sum,
                long long *nan_count, void *stream) {
    if (!adam) { set_error("null adam descriptor"); return SVBASL_E_INVALID; }
    return run_step(model, engine, adam, 0, nullptr, nullptr, cost_sum, nan_count, stream);
}

int svbasl_step_spatial(const svbasl_model *model, const svbasl_engine *engine, const svbasl_adam *adam,
                        const svbasl_hyper *hyper, double *cost_sum, long long *nan_count, void *stream) {
    if (!adam) { set_error("null adam descriptor"); return SVBASL_E_INVALID; }
    return run_step(model, engine, adam, 0, nullptr, nullptr, cost_sum, nan_count, stream, hyper);
}

static int launch_spatial_samples(const svbasl_engine *engine, int64_t first, int64_t count, int64_t step, float *out,
                                  bool mirror, void *stream) {
    if (count == 0 || mrf_mask(engine) == 0) return 0;
    SpatialArgs a;
    a.e = *engine;
    if (!mirror) a.e.peer_lo = a.e.peer_hi = nullptr;
    a.ec = make_engine_const(*engine);
    a.first = first;
    a.count = count;
    a.step = step;
    a.out = out;
    spatial_sample_kernel<<<(unsigned)((count + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(a);
    return check_launch("spatial_sample_kernel");
}

int svbasl_sample_spatial(const svbasl_engine *engine, int64_t n_local, int64_t step, float *out, void *stream) {
    if (!engine || !out || !engine->state || n_local < 0 || n_local > engine->ld || engine->n_par < 1 ||
        engine->n_par > SVBASL_MAX_PAR || engine->n_samples < 1) {
        set_error("bad sample_spatial arguments");
        return SVBASL_E_INVALID;
    }
    return launch_spatial_samples(engine, 0, n_local, step, out, false, stream);
}

int svbasl_sample_spatial_next(const svbasl_engine *engine, int64_t step, float *out, void *stream) {
    if (!engine || !out || !engine->state || engine->n_vox < 0 || engine->w_begin < 0 ||
        engine->w_begin + engine->n_vox > engine->ld || engine->n_par < 1 || engine->n_par > SVBASL_MAX_PAR ||
        engine->n_samples < 1 || engine->eps) {
        set_error("bad sample_spatial_next arguments (needs the in-kernel draws)");
        return SVBASL_E_INVALID;
    }
    if (engine->peer_lo || engine->peer_hi) {
        const int64_t nlo = engine->peer_lo ? engine->peer_lo_count : 0, nhi = engine->peer_hi ? engine->peer_hi_count : 0;
        if (nlo < 0 || nhi < 0 || nlo + nhi > engine->n_vox || (engine->peer_lo && engine->peer_lo_first != engine->w_begin) ||
            (engine->peer_hi && engine->peer_hi_first != engine->w_begin + engine->n_vox - nhi)) {
            set_error("peer_lo / peer_hi must name the first / last owned voxels");
            return SVBASL_E_INVALID;
        }
    }
    return launch_spatial_samples(engine, engine->w_begin, engine->n_vox, step, out, true, stream);
}

int svbasl_hyper_step(float *log_ak, float *m, float *v, const double *ak_grad, int32_t n, float grad_scale, float lr_t,
                      float beta1, float beta2, float epsilon, void *stream) {
    if (!log_ak || !m || !v || !ak_grad || n < 1 || n > SVBASL_MAX_SPATIAL) { set_error("bad hyper_step arguments"); return SVBASL_E_INVALID; }
    hyper_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(log_ak, m, v, ak_grad, n, grad_scale, lr_t, beta1, beta2, epsilon);
    return check_launch("hyper_step_kernel");
}

int svbasl_hyper_step_dev(float *log_ak, float *m, float *v, double *ak_grad, int32_t n, float grad_scale,
                          const float *lr_t, long long *step_dev, float beta1, float beta2, float epsilon, void *stream) {
    if (!log_ak || !m || !v || !ak_grad || !lr_t || !step_dev || n < 1 || n > SVBASL_MAX_SPATIAL) {
        set_error("bad hyper_step_dev arguments");
        return SVBASL_E_INVALID;
    }
    hyper_step_dev_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(log_ak, m, v, ak_grad, n, grad_scale, lr_t, step_dev, beta1,
                                                              beta2, epsilon);
    return check_launch("hyper_step_dev_kernel");
}

int64_t svbasl_mailbox_bytes(int32_t world) { return world > 0 ? (int64_t)sizeof(MailSlot) * 2 * world : 0; }

int svbasl_hyper_step_peers(float *log_ak, float *m, float *v, double *ak_grad, int32_t n, float grad_scale,
                            const float *lr_t, long long *step_dev, float beta1, float beta2, float epsilon,
                            int32_t rank, int32_t world, void *const *mailboxes, int32_t *status, void *stream) {
    if (!log_ak || !m || !v || !ak_grad || !lr_t || !step_dev || !mailboxes || !status || n < 1 || n > SVBASL_MAX_SPATIAL ||
        world < 1 || world > SVBASL_MAX_PEERS || rank < 0 || rank >= world) {
        set_error("bad hyper_step_peers arguments");
        return SVBASL_E_INVALID;
    }
    svbasl_hyper hy;
    memset(&hy, 0, sizeof(hy));
    for (int r = 0; r < world; ++r) {
        if (!mailboxes[r]) { set_error("mailbox of rank %d is NULL", r); return SVBASL_E_INVALID; }
        hy.mailboxes[r] = mailboxes[r];
    }
    hy.log_ak = log_ak; hy.m = m; hy.v = v; hy.lr_t = lr_t; hy.step_dev = step_dev;
    hy.beta1 = beta1; hy.beta2 = beta2; hy.epsilon = epsilon;
    hy.n_spatial = n; hy.rank = rank; hy.world = world; hy.status = status;
    hyper_step_peers_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hy, ak_grad, grad_scale);
    return check_launch("hyper_step_peers_kernel");
}

int svbasl_shared_alloc(int64_t bytes, void **dev_ptr, unsigned char handle[SVBASL_IPC_HANDLE_BYTES]) {
    if (bytes <= 0 || !dev_ptr || !handle) { set_error("bad shared_alloc arguments"); return SVBASL_E_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == SVBASL_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
    CUDA_TRY(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, p);
    if (err != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(err));
        return SVBASL_E_CUDA;
    }
    memcpy(handle, &h, sizeof(h));
    *dev_ptr = p;
    return 0;
}

int svbasl_shared_free(void *dev_ptr) {
    if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
    return 0;
}

int svbasl_shared_open(const unsigned char handle[SVBASL_IPC_HANDLE_BYTES], void **dev_ptr) {
    if (!handle || !dev_ptr) { set_error("bad shared_open arguments"); return SVBASL_E_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // current device = the accessing device
    *dev_ptr = p;
    return 0;
}

int svbasl_shared_close(void *dev_ptr) {
    if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

int svbasl_advance_step(long long *step_dev, long long inc, void *stream) {
    if (!step_dev) { set_error("null step counter"); return SVBASL_E_INVALID; }
    advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, inc);
    return check_launch("advance_step_kernel");
}

int svbasl_fill_eps(float *eps, int64_t n_vox, int64_t ld, int64_t vox_offset, int32_t n_par, int32_t n_samples,
                    uint64_t seed, int64_t step, void *stream) {
    if (!eps || n_vox < 0 || ld < n_vox || n_par < 1 || n_samples < 1) { set_error("bad fill_eps arguments"); return SVBASL_E_INVALID; }
    if (n_vox == 0) return 0;
    fill_eps_kernel<<<(unsigned)((n_vox + 127) / 128), 128, 0, (cudaStream_t)stream>>>(eps, n_vox, ld, vox_offset, n_par,
                                                                                      n_samples, seed, step);
    return check_launch("fill_eps_kernel");
}

int svbasl_init_stats(const float *data, const float *tpts, int64_t n_vox, int64_t ld, int32_t t_full, float *mean_t,
                      float *max_t, float *var_t, float *t_at_max, void *stream) {
    if (!data || n_vox < 0 || ld < n_vox || t_full < 1) { set_error("bad init_stats arguments"); return SVBASL_E_INVALID; }
    if (n_vox == 0) return 0;
    init_stats_kernel<<<(unsigned)((n_vox + 127) / 128), 128, 0, (cudaStream_t)stream>>>(data, tpts, n_vox, ld, t_full, mean_t,
                                                                                        max_t, var_t, t_at_max);
    return check_launch("init_stats_kernel");
}

int svbasl_model_fit(const svbasl_model *model, const svbasl_engine *engine, float *out, void *stream) {
    if (!model || !engine || !out) { set_error("null argument"); return SVBASL_E_INVALID; }
    const KernelEntry *k = find_entry(model, 0, true);
    if (!k) { set_error("no kernel compiled for model kind=%d flags=0x%x", model->kind, canonical_flags(model)); return SVBASL_E_UNSUPPORTED; }
    if (engine->n_par != k->n_params + 1 || !engine->state || (!engine->tpts && !engine->ti)) {
        set_error("bad engine descriptor for model_fit");
        return SVBASL_E_INVALID;
    }
    FitArgs a;
    a.md = make_dev_model(*model);
    a.e = *engine;
    a.out = out;
    return k->fit(a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// Host-staged iteration: the reference feeds every batch from host memory (sess.run(feed_dict)); here
// the batch of iteration i+1 is copied on a second stream while iteration i computes.
int svbasl_host_ctx_create(svbasl_host_ctx **out, int64_t ld, int32_t n_batch) {
    if (!out || ld < 1 || n_batch < 1) { set_error("bad host_ctx arguments"); return SVBASL_E_INVALID; }
    svbasl_host_ctx *c = new svbasl_host_ctx();
    memset(c, 0, sizeof(*c));
    c->ld = ld;
    c->n_batch = n_batch;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->run_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&c->copied[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->consumed[i], cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&c->d_data[i], sizeof(float) * ld * n_batch));
        CUDA_TRY(cudaMalloc(&c->d_tpts[i], sizeof(float) * ld * n_batch));
        CUDA_TRY(cudaMalloc(&c->d_ti[i], sizeof(float) * n_batch));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&c->d_cost, sizeof(double) * 2));
    *out = c;
    return 0;
}

int svbasl_host_ctx_destroy(svbasl_host_ctx *c) {
    if (!c) return 0;
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->run_stream);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_data[i]);
        cudaFree(c->d_tpts[i]);
        cudaFree(c->d_ti[i]);
        cudaEventDestroy(c->copied[i]);
        cudaEventDestroy(c->consumed[i]);
    }
    cudaFree(c->d_cost);
    cudaEventDestroy(c->done);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->run_stream);
    delete c;
    return 0;
}

int svbasl_step_host(svbasl_host_ctx *c, const svbasl_model *model, const svbasl_engine *engine, const svbasl_adam *adam,
                     const svbasl_hyper *hyper, const float *host_data, const float *host_tpts, const float *host_ti,
                     double *host_cost_sum) {
    if (!c || !engine || !adam || !host_data || (!host_tpts && !host_ti)) { set_error("null argument"); return SVBASL_E_INVALID; }
    if (engine->ld != c->ld || engine->n_batch != c->n_batch) { set_error("engine does not match the host context"); return SVBASL_E_INVALID; }
    const int s = c->slot;
    const size_t bytes = sizeof(float) * (size_t)c->ld * (size_t)c->n_batch;
    // the staging slot may only be overwritten once the iteration that used it two calls ago has finished
    if (c->calls >= 2) CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->consumed[s], 0));
    CUDA_TRY(cudaMemcpyAsync(c->d_data[s], host_data, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    if (host_tpts) CUDA_TRY(cudaMemcpyAsync(c->d_tpts[s], host_tpts, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    else CUDA_TRY(cudaMemcpyAsync(c->d_ti[s], host_ti, sizeof(float) * (size_t)c->n_batch, cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->copied[s], c->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(c->run_stream, c->copied[s], 0));
    svbasl_engine e = *engine;
    e.data = c->d_data[s];
    e.tpts = host_tpts ? c->d_tpts[s] : nullptr;
    if (!host_tpts) e.ti = c->d_ti[s];            // e.zoff stays the caller's resident per-voxel slice offset
    e.t_row0 = 0;
    e.t_row_stride = 1;
    e.cost_sum_scalar = 1;
    svbasl_adam ad = *adam;
    ad.n_iters = 1;
    ad.n_batches = 1;
    double *d_cost = host_cost_sum ? c->d_cost + s : nullptr;
    if (d_cost) CUDA_TRY(cudaMemsetAsync(d_cost, 0, sizeof(double), c->run_stream));
    int rc = svbasl_step_spatial(model, &e, &ad, hyper, d_cost, nullptr, c->run_stream);
    if (rc) return rc;
    if (d_cost) CUDA_TRY(cudaMemcpyAsync(host_cost_sum, d_cost, sizeof(double), cudaMemcpyDeviceToHost, c->run_stream));
    CUDA_TRY(cudaEventRecord(c->consumed[s], c->run_stream));
    c->slot ^= 1;
    c->calls++;
    return 0;
}

int svbasl_host_sync(svbasl_host_ctx *c) {
    if (!c) { set_error("null context"); return SVBASL_E_INVALID; }
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(c->run_stream));
    return 0;
}

}  // extern "C"
