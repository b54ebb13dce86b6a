// voxel_step.h - one voxel's share of an SVB iteration, entirely in registers:
//   reparameterised samples theta_s = mu + L eps_s from the per-voxel MVN posterior,
//   model prediction + analytic model derivatives per (sample, time point),
//   Gaussian-noise negative log-likelihood, latent loss (sample-based or closed-form KL) against
//   N / ARD / spatial-MRF priors, the hand-derived gradient with respect to (mu, log-variance,
//   Cholesky off-diagonals, ARD log-precision), and optionally the TensorFlow-form Adam update.
//
// This is the part of the external svb engine (SvbFit graph: posterior.sample -> model.evaluate ->
// noise.log_likelihood -> latent loss -> tf.gradients -> AdamOptimizer)
// that sits on the hot path around /root/reference/svb_models_asl/aslrest.py:248-340.  Backward-pass algebra:
// SURVEY.md Appendix A.3 / DESIGN.md section 3.  One thread owns one voxel; nothing of size [W,S,B] ever exists.
#pragma once
#include <type_traits>
#include "compat.h"
#include "philox.h"
#include "dev_model.h"

#ifndef SVB_PAIRED_SAMPLES
#define SVB_PAIRED_SAMPLES 1
#endif

namespace svb {

SVB_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }       // lower triangle incl. diagonal
SVB_HD constexpr int stri(int i, int j) { return i * (i - 1) / 2 + j; }      // strict lower triangle (state order)

// Engine-level constants folded on the host (no divisions / logs per voxel)
struct EngineConst {
    float pinv[SVBASL_MAX_PAR];     // 1 / prior variance
    float plog[SVBASL_MAX_PAR];     // log prior variance
    float t_full, scale, inv_s;     // T, T/B, 1/S
    int32_t sp_slot[SVBASL_MAX_PAR];// index of parameter i among the spatial ("M") parameters, or -1
};

inline EngineConst make_engine_const(const svbasl_engine &e) {
    EngineConst c;
    int n_sp = 0;
    for (int i = 0; i < SVBASL_MAX_PAR; ++i) {
        const bool ok = i < e.n_par && e.prior_var[i] > 0.0f;
        c.pinv[i] = ok ? 1.0f / e.prior_var[i] : 0.0f;
        c.plog[i] = ok ? logf(e.prior_var[i]) : 0.0f;
        c.sp_slot[i] = (i < e.n_par && e.prior_type[i] == SVBASL_PRIOR_MRF) ? n_sp++ : -1;
    }
    c.t_full = (float)e.t_full;
    c.scale = (float)e.t_full / (float)e.n_batch;
    c.inv_s = 1.0f / (float)e.n_samples;
    return c;
}

// Residual accumulator handed to Model::run(): keeps the batch's data/time points in registers (NBT > 0)
// or re-reads them through L1 (NBT == 0, any batch size).
template <int P, int NBT>
struct BatchAcc {
    static constexpr int NB = NBT;
    float y[NBT > 0 ? NBT : 1], t[NBT > 0 ? NBT : 1];
    const float *yp, *tp, *tip;     // dynamic mode: voxel-column bases
    int64_t stride;                 // floats between consecutive batch rows
    int ti_stride;
    float zoff;
    int nb;
    float ssd;
    float G[P > 0 ? P : 1];

    SVB_HD void load(const svbasl_engine &e, int64_t w, int row0) {
        nb = e.n_batch;
        stride = (int64_t)e.t_row_stride * e.ld;
        yp = e.data + (int64_t)row0 * e.ld + w;
        tp = e.tpts ? e.tpts + (int64_t)row0 * e.ld + w : nullptr;
        tip = e.ti ? e.ti + row0 : nullptr;
        ti_stride = e.t_row_stride;
        zoff = (e.zoff && !e.tpts) ? e.zoff[w] : 0.0f;
        if (NBT > 0) {
#pragma unroll
            for (int b = 0; b < (NBT > 0 ? NBT : 1); ++b) {
                y[b] = yp[b * stride];
                t[b] = tp ? tp[b * stride] : tip[b * ti_stride] + zoff;
            }
        }
    }
    SVB_HD int n() const { return nb; }
    SVB_HD float time(int b) const {
        if (NBT > 0) return t[b];
        return tp ? tp[b * stride] : tip[b * ti_stride] + zoff;
    }
    SVB_HD void reset() {
        ssd = 0.0f;
#pragma unroll
        for (int p = 0; p < P; ++p) G[p] = 0.0f;
    }
    // residual of time point b, accumulated into the sum of squares; the model forms its own derivative sums
    SVB_HD float resid(int b, float pred) {
        const float r = pred - (NBT > 0 ? y[b] : yp[b * stride]);
        ssd += r * r;
        return r;
    }
    SVB_HD void add(int b, float pred, const float *d) {
        float r = pred - (NBT > 0 ? y[b] : yp[b * stride]);
        ssd += r * r;
#pragma unroll
        for (int p = 0; p < P; ++p) G[p] += r * d[p];
    }
};

// One posterior sample theta_p (internal value) of voxel u for sample s: the quantity the spatial prior
// compares between neighbours.  Written by the pre-pass (kernels.cuh: spatial_sample_kernel) into
// e.spatial_samples so that the main kernel reads its neighbours' samples instead of rebuilding them.
// n = P' (runtime here: the pre-pass is model independent).
SVB_HD float sample_theta(const svbasl_engine &e, uint32_t key, int64_t u, int p, int s) {
    const int n = e.n_par;
    const float *st = e.state + u;
    float th = st[(int64_t)p * e.ld];
    for (int j = 0; j <= p; ++j) {
        const float ej = e.eps ? e.eps[((int64_t)j * e.n_samples + s) * e.ld + u]
                               : normal_at(key, e.vox_offset + u, j, s, e.n_samples);
        th += (j == p ? fexp(0.5f * st[(int64_t)(n + p) * e.ld]) : st[(int64_t)(2 * n + stri(p, j)) * e.ld]) * ej;
    }
    return th;
}

// The neighbours' samples of ONE spatial parameter, staged per thread (kernels.cuh copies them into shared memory
// asynchronously at kernel start): v[(s*6 + k) * stride] = sample s of neighbour k; a neighbour that does not exist is
// represented by the voxel's own sample (zero difference).
// v == nullptr: no tile, elbo_grad gathers from e.spatial_samples where it needs them.
struct NbTile {
    const float *v;
    int stride;
    int param;
};

template <class M, class = void>
struct paired_samples : std::false_type {};
template <class M>
struct paired_samples<M, std::void_t<decltype(M::kPairedSamples)>> : std::bool_constant<M::kPairedSamples> {};

// FL, the flavour of a step: 0 = generic (every run-time switch live); 1 = lean, the production flavour -
// sample-based latent loss, draws from the in-register Philox stream, no spatial prior - with those three
// switches resolved at compile time, so the hot loop carries no dead code (the generic flavour's skipped
// branches cost instruction-cache misses, profiles/r1_notes.md); 2 = lean with the spatial prior kept.
template <class M, int NBT, int FL = 0>
struct VoxelStep {
    static constexpr bool LEAN = FL != 0;           // numeric latent loss + Philox draws fixed at compile time
    static constexpr bool SPATIAL = FL != 1;
    // sample loop unrolled over the two samples of a Philox call: for the models that ask for it, in the production
    // flavour without the spatial prior (aslrest: +4.8 %; the spatial flavour is 7 % faster with the rolled loop, and the
    // tensor-core aslnn step loses a quarter of its rate with two copies of its pipelined row loop, profiles/r3_notes.md)
    static constexpr bool PAIRED = SVB_PAIRED_SAMPLES && paired_samples<M>::value && FL == 1;
    static constexpr int P = M::P;
    static constexpr int N = P + 1;                 // noise last
    static constexpr int NL = N * (N - 1) / 2;
    static constexpr int NT = N * (N + 1) / 2;

    // posterior state of this voxel
    float mu[N], lv[N], od[NL > 0 ? NL : 1];
    float lphi[N];                                  // ARD log phi (only ARD slots used)
    // gradients of grad_scale*cost (filled by elbo_grad)
    float g_mu[N], g_lv[N], g_od[NL > 0 ? NL : 1], g_lphi[N];
    float ak_out[N];                                // share of d(sum cost)/d(log ak) of spatial parameter i

    SVB_HD void load(const svbasl_engine &e, int64_t w) { load_rows(e, e.state + w, e.ld); }

    // state rows of one voxel from `s` with row stride `ld` (global memory, or a staged copy)
    SVB_HD void load_rows(const svbasl_engine &e, const float *s, int64_t ld) {
        int n_ard = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            mu[i] = s[(int64_t)i * ld];
            lv[i] = s[(int64_t)(N + i) * ld];
            lphi[i] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < NL; ++k) od[k] = s[(int64_t)(2 * N + k) * ld];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (e.prior_type[i] == SVBASL_PRIOR_ARD) {
                lphi[i] = s[(int64_t)(2 * N + NL + n_ard) * ld];
                ++n_ard;
            }
        }
    }

    // Per-voxel prior / posterior terms that the sample loop and the closing algebra share
    struct Terms {
        float sd[N];                                   // exp(logvar / 2)
        float pm[N], pinv[N], plog[N], phi_live[N];    // prior mean, 1/variance, log variance; ARD gradient gate
        float lw_pinv[N];                              // latent_weight / prior variance
    };

    SVB_HD void prior_terms(const svbasl_engine &e, const EngineConst &ec, Terms &t) const {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            t.sd[i] = fexp(0.5f * lv[i]);
            t.pm[i] = e.prior_mean[i];
            if (e.prior_type[i] == SVBASL_PRIOR_ARD) {
                float phi = fexp(lphi[i]);
                bool clipped = (e.ard_phi_max > 0.0f) && (phi > e.ard_phi_max);
                phi = clipped ? e.ard_phi_max : phi;
                t.pinv[i] = phi;
                t.plog[i] = -flog(phi);
                t.phi_live[i] = clipped ? 0.0f : 1.0f;
            } else {
                t.pinv[i] = ec.pinv[i];
                t.plog[i] = ec.plog[i];
                t.phi_live[i] = 0.0f;
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) t.lw_pinv[i] = e.latent_weight * t.pinv[i];
    }

    // Closing algebra of elbo_grad: from the sums over samples (a_mu = sum_s g_s, a_L = sum_s g_s eps_s^T,
    // a_hyp = sum_s (theta - m)^2 or the MRF log-ak share, cost = sum_s of the per-sample cost) to the cost of
    // the voxel and the gradients g_* (latent-loss constants, 1/S, entropy or closed-form KL, chain rule to logvar).
    SVB_HD float finish(const svbasl_engine &e, const EngineConst &ec, const Terms &t, float *a_mu, float *a_L, float *a_hyp,
                        float cost, bool numeric) {
        const int S = e.n_samples;
        const float lw = e.latent_weight;
        const float invS = ec.inv_s;
        const float gs = e.grad_scale;
        const float *sd = t.sd, *pm = t.pm, *pinv = t.pinv, *plog = t.plog, *phi_live = t.phi_live;
        if (numeric) {
            // Sample-based latent loss: every gradient is (1/S) x (sum over samples) plus, on the diagonal of L, the
            // entropy term -1/2 log det(cov) = -sum_i log L_ii.  With L_ii = exp(logvar_i / 2) the chain rule turns its
            // gradient -lw / L_ii into the constant -lw/2 per log-variance, so 1/S and grad_scale are applied as ONE
            // factor to the raw sums.
            const float kg = invS * gs;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (!SPATIAL || e.prior_type[i] != SVBASL_PRIOR_MRF) {
                    const float q = t.lw_pinv[i] * a_hyp[i];            // sum_s lw (theta-m)^2 / v
                    cost += 0.5f * q + (float)S * lw * 0.5f * plog[i];
                    a_hyp[i] = 0.5f * (q - (float)S * lw);              // sum_s lw/2 ((theta-m)^2/v - 1)
                }
            }
            cost *= invS;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cost -= lw * 0.5f * lv[i];
                g_mu[i] = kg * a_mu[i];
                g_lv[i] = (kg * a_L[tri(i, i)]) * (0.5f * sd[i]) - 0.5f * gs * lw;
                g_lphi[i] = kg * a_hyp[i] * phi_live[i];
                if (SPATIAL) ak_out[i] = invS * a_hyp[i];
#pragma unroll
                for (int j = 0; j < i; ++j) g_od[stri(i, j)] = kg * a_L[tri(i, j)];
            }
            return cost;
        }
        cost *= invS;
#pragma unroll
        for (int i = 0; i < N; ++i) { a_mu[i] *= invS; a_hyp[i] *= invS; }
#pragma unroll
        for (int k = 0; k < NT; ++k) a_L[k] *= invS;
        // closed-form KL( N(mu, cov) || N(pm, diag(pv)) ), cov = L^T L (svb) or L L^T
        float kl = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float dm = mu[i] - pm[i];
            float cii = 0.0f;     // cov_ii
            if (e.cov_llt) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    const float l = (j == i) ? sd[i] : od[stri(i, j)];
                    cii += l * l;
                    a_L[tri(i, j)] += lw * l * pinv[i];
                }
            } else {
#pragma unroll
                for (int r = i; r < N; ++r) {
                    const float l = (r == i) ? sd[i] : od[stri(r, i)];
                    cii += l * l;
                    a_L[tri(r, i)] += lw * l * pinv[i];
                }
            }
            kl += cii * pinv[i] + dm * dm * pinv[i] - 1.0f + plog[i] - lv[i];
            a_mu[i] += lw * dm * pinv[i];
            a_L[tri(i, i)] -= lw * frcp(sd[i]);
            a_hyp[i] = lw * 0.5f * (pinv[i] * (cii + dm * dm) - 1.0f);
        }
        cost += lw * 0.5f * kl;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            g_mu[i] = gs * a_mu[i];
            g_lv[i] = gs * a_L[tri(i, i)] * 0.5f * sd[i];
            g_lphi[i] = gs * a_hyp[i] * phi_live[i];
            ak_out[i] = a_hyp[i];
#pragma unroll
            for (int j = 0; j < i; ++j) g_od[stri(i, j)] = gs * a_L[tri(i, j)];
        }
        return cost;
    }

    // Cost of this voxel for one batch, gradients left in g_*.  Returns the un-scaled cost.
    SVB_HD float elbo_grad(const DevModel &md, const svbasl_engine &e, const EngineConst &ec, int64_t w, int64_t step,
                           int row0, const NbTile nbt = NbTile{nullptr, 0, -1}) {
        typename M::Vox vox = M::load_vox(md, w);
        BatchAcc<P, NBT> acc;
        acc.load(e, w, row0);
        M::bind_times(md, vox, acc);
        return elbo_grad_batch(md, e, ec, w, step, vox, acc, nbt);
    }

    // The same with the batch (data, time points, per-time-point model constants) already loaded: the iterations fused
    // into one launch share it when every iteration sees the same batch (n_batches == 1).
    SVB_HD float elbo_grad_batch(const DevModel &md, const svbasl_engine &e, const EngineConst &ec, int64_t w, int64_t step,
                                 typename M::Vox &vox, BatchAcc<P, NBT> &acc, const NbTile nbt = NbTile{nullptr, 0, -1}) {
        const int S = e.n_samples;
        const float Tf = ec.t_full, half_Tf = 0.5f * ec.t_full;
        const float scale = ec.scale;
        const float lw = e.latent_weight;
        const bool numeric = LEAN || (e.latent == SVBASL_LATENT_NUMERIC);
        const bool eps_mem = !LEAN && e.eps != nullptr;
        const uint32_t key = rng_key(e.seed, step);

        Terms tm;
        prior_terms(e, ec, tm);
        const float *sd = tm.sd, *pm = tm.pm, *lw_pinv = tm.lw_pinv;
        // a_hyp[i]: ARD log-phi gradient (ARD parameters) or log-ak gradient share (spatial parameters)
        float a_mu[N], a_L[NT], a_hyp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { a_mu[i] = 0.0f; a_hyp[i] = 0.0f; }
#pragma unroll
        for (int k = 0; k < NT; ++k) a_L[k] = 0.0f;
        float cost = 0.0f;

        // One sample: theta = mu + L eps, the model and its derivatives over the batch, the per-sample cost and the
        // sums the closing algebra needs.
        auto one_sample = [&](int s, const float *eps) {
            float th[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float v = mu[i] + sd[i] * eps[i];
#pragma unroll
                for (int j = 0; j < i; ++j) v += od[stri(i, j)] * eps[j];
                th[i] = v;
            }
            float x[P > 0 ? P : 1], dx[P > 0 ? P : 1];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int code = M::xf(p);
                if (code == SVBASL_XF_EXP) { x[p] = fexp(th[p]); dx[p] = x[p]; }
                else if (code == SVBASL_XF_ABS) { x[p] = fabsf(th[p]); dx[p] = th[p] < 0.0f ? -1.0f : 1.0f; }
                else { x[p] = th[p]; dx[p] = 1.0f; }
            }
            acc.reset();
            M::run(md, vox, x, acc);

            const float thn = th[N - 1];
            const float inv_nv = fexp(-thn);
            const float sc = scale * inv_nv;
            const float c_ssd = sc * acc.ssd;
            cost += 0.5f * (Tf * thn + c_ssd);
            float g[N];
#pragma unroll
            for (int p = 0; p < P; ++p) g[p] = sc * acc.G[p] * dx[p];
            g[N - 1] = half_Tf - 0.5f * c_ssd;
            if (numeric) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    if (SPATIAL && e.prior_type[i] == SVBASL_PRIOR_MRF) {
                        // -E_s[ 1/2 log ak - ak/4 sum_u (x_w - x_u)^2 ]  (SURVEY Appendix A.5); the neighbours'
                        // samples come from the pre-pass buffer [slot][S][ld]
                        const int slot = ec.sp_slot[i];
                        const float lak = e.log_ak[slot];
                        const float ak = fexp(lak);
                        const float *nbs = e.spatial_samples + ((int64_t)slot * S + s) * e.ld;
                        float sdx = 0.0f, sdx2 = 0.0f;
                        if (nbt.v && nbt.param == i) {
                            if (s == 0) async_copies_wait();
                            const float *nv = nbt.v + (int64_t)s * 6 * nbt.stride;
#pragma unroll
                            for (int nbr = 0; nbr < 6; ++nbr) {
                                const float dxu = th[i] - nv[nbr * nbt.stride];
                                sdx += dxu;
                                sdx2 += dxu * dxu;
                            }
                        } else {
#pragma unroll
                            for (int nbr = 0; nbr < 6; ++nbr) {
                                const int u = e.neighbours[(int64_t)nbr * e.ld + w];
                                if (u >= 0) {
                                    const float dxu = th[i] - nbs[u];
                                    sdx += dxu;
                                    sdx2 += dxu * dxu;
                                }
                            }
                        }
                        cost += lw * (-0.5f * lak + 0.25f * ak * sdx2);
                        g[i] += lw * ak * sdx;          // own term + the symmetric term of each neighbour's cost
                        a_hyp[i] += lw * (-0.5f + 0.25f * ak * sdx2);
                    } else {
                        // -log N(theta; m, v) up to the constant 1/2 log v (added once after the loop);
                        // a_hyp accumulates (theta-m)^2; times lw/v (finish) it is also the ARD log-phi gradient term
                        const float dth = th[i] - pm[i];
                        g[i] += dth * lw_pinv[i];
                        a_hyp[i] += dth * dth;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                a_mu[i] += g[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) a_L[tri(i, j)] += g[i] * eps[j];
            }
        };
        if (eps_mem) {
            for (int s = 0; s < S; ++s) {
                float eps[N];
#pragma unroll
                for (int j = 0; j < N; ++j) eps[j] = e.eps[((int64_t)j * S + s) * e.ld + w];
                one_sample(s, eps);
            }
        } else if (PAIRED) {
            // a Philox call yields the same posterior row of two consecutive samples (philox.h): draw both, run both
            for (int s = 0; s < S; s += 2) {
                float ea[N], eb[N];
#pragma unroll
                for (int j = 0; j < N; ++j) normal_pair(key, e.vox_offset + w, stream_pair(j, s, S), ea[j], eb[j]);
                one_sample(s, ea);
                if (s + 1 < S) one_sample(s + 1, eb);
            }
        } else {
            float spare[N];                            // second halves of the Box-Muller pairs: the next sample's draws
#pragma unroll
            for (int j = 0; j < N; ++j) spare[j] = 0.0f;
            for (int s = 0; s < S; ++s) {
                float eps[N];
                normal_row<N>(key, e.vox_offset + w, s, S, eps, spare);
                one_sample(s, eps);
            }
        }
        return finish(e, ec, tm, a_mu, a_L, a_hyp, cost, numeric);
    }

    SVB_HD bool grads_finite() const {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) acc += g_mu[i] * 0.0f + g_lv[i] * 0.0f + g_lphi[i] * 0.0f;
#pragma unroll
        for (int k = 0; k < NL; ++k) acc += g_od[k] * 0.0f;
        return acc == 0.0f;               // NaN/Inf * 0 = NaN
    }

    SVB_HD void store_grads(const svbasl_engine &e, float *grad, int64_t w) const { store_grad_rows(e, grad + w, e.ld); }

    SVB_HD void store_grad_rows(const svbasl_engine &e, float *g, int64_t ld) const {
        int a = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            g[(int64_t)i * ld] = g_mu[i];
            g[(int64_t)(N + i) * ld] = g_lv[i];
        }
#pragma unroll
        for (int k = 0; k < NL; ++k) g[(int64_t)(2 * N + k) * ld] = g_od[k];
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (e.prior_type[i] == SVBASL_PRIOR_ARD) g[(int64_t)(2 * N + NL + a++) * ld] = g_lphi[i];
    }

    static SVB_HD float adam1(const svbasl_adam &ad, float lr_t, float x, float g, float &m, float &v) {
        m += (g - m) * (1.0f - ad.beta1);              // the form TensorFlow's ApplyAdam kernel evaluates
        v += (g * g - v) * (1.0f - ad.beta2);
        return x - lr_t * fdiv(m, fsqrt(v) + ad.epsilon);
    }

    // tf.train.AdamOptimizer step on this voxel's rows.  The old moments are read from (m_rd, v_rd) with row
    // stride rd_stride and the new ones written to (m_wr, v_wr) with row stride wr_stride - global memory or the
    // shared-memory tile the kernel prefetched them into (between the fused iterations of one launch the moments
    // never leave shared memory).  With write_state the new state is stored too.
    SVB_HD void adam_update(const svbasl_engine &e, const svbasl_adam &ad, float lr_t, int64_t w, bool write_state,
                            const float *m_rd, const float *v_rd, int64_t rd_stride, float *m_wr, float *v_wr,
                            int64_t wr_stride) {
        float *s = (e.state_out ? e.state_out : e.state) + w;
        int a = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float mm = m_rd[(int64_t)i * rd_stride], vv = v_rd[(int64_t)i * rd_stride];
            mu[i] = adam1(ad, lr_t, mu[i], g_mu[i], mm, vv);
            m_wr[(int64_t)i * wr_stride] = mm;
            v_wr[(int64_t)i * wr_stride] = vv;
            if (write_state) s[(int64_t)i * e.ld] = mu[i];
            mm = m_rd[(int64_t)(N + i) * rd_stride]; vv = v_rd[(int64_t)(N + i) * rd_stride];
            lv[i] = adam1(ad, lr_t, lv[i], g_lv[i], mm, vv);
            m_wr[(int64_t)(N + i) * wr_stride] = mm;
            v_wr[(int64_t)(N + i) * wr_stride] = vv;
            if (write_state) s[(int64_t)(N + i) * e.ld] = lv[i];
        }
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            float mm = m_rd[(int64_t)(2 * N + k) * rd_stride], vv = v_rd[(int64_t)(2 * N + k) * rd_stride];
            od[k] = adam1(ad, lr_t, od[k], g_od[k], mm, vv);
            m_wr[(int64_t)(2 * N + k) * wr_stride] = mm;
            v_wr[(int64_t)(2 * N + k) * wr_stride] = vv;
            if (write_state) s[(int64_t)(2 * N + k) * e.ld] = od[k];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (e.prior_type[i] == SVBASL_PRIOR_ARD) {
                const int row = 2 * N + NL + a++;
                float mm = m_rd[(int64_t)row * rd_stride], vv = v_rd[(int64_t)row * rd_stride];
                lphi[i] = adam1(ad, lr_t, lphi[i], g_lphi[i], mm, vv);
                m_wr[(int64_t)row * wr_stride] = mm;
                v_wr[(int64_t)row * wr_stride] = vv;
                if (write_state) s[(int64_t)row * e.ld] = lphi[i];
            }
        }
    }

    // Samples of the NEXT iteration for the spatial prior, from the state this step has just produced:
    // theta_{p,s}(step+1) = mu_p + sum_j L_pj eps_j,s with the draws of step+1 (the same arithmetic, term by term, as
    // the pre-pass sample_theta / spatial_sample_kernel), stored to e.spatial_samples_out [slot][S][ld] - the buffer the
    // neighbours read in the next launch - and, for a shard-boundary voxel, ALSO into the adjacent rank's halo column
    // through NVLink peer memory (svbasl_engine.peer_*): the halo "exchange" is these stores, no launch of its own.
    SVB_HD void store_next_samples(const svbasl_engine &e, const EngineConst &ec, int64_t w, int64_t next_step) const {
        int pmax = -1;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (e.prior_type[i] == SVBASL_PRIOR_MRF) pmax = i;
        if (pmax < 0) return;
        const uint32_t key = rng_key(e.seed, next_step);
        const int S = e.n_samples;
        float sdn[N];
#pragma unroll
        for (int i = 0; i < N; ++i) sdn[i] = (e.prior_type[i] == SVBASL_PRIOR_MRF) ? fexp(0.5f * lv[i]) : 0.0f;
        float *own = e.spatial_samples_out + w;
        float *lo = (e.peer_lo && w >= e.peer_lo_first && w < e.peer_lo_first + e.peer_lo_count)
                        ? e.peer_lo + (w + e.peer_lo_shift) : nullptr;
        float *hi = (e.peer_hi && w >= e.peer_hi_first && w < e.peer_hi_first + e.peer_hi_count)
                        ? e.peer_hi + (w + e.peer_hi_shift) : nullptr;
        for (int s0 = 0; s0 < S; s0 += 2) {                  // two samples per Philox call and row
            float ea[N], eb[N];
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (j <= pmax) normal_pair(key, e.vox_offset + w, stream_pair(j, s0, S), ea[j], eb[j]);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (e.prior_type[i] != SVBASL_PRIOR_MRF) continue;
                float tha = mu[i], thb = mu[i];
#pragma unroll
                for (int j = 0; j < i; ++j) { tha += od[stri(i, j)] * ea[j]; thb += od[stri(i, j)] * eb[j]; }
                tha += sdn[i] * ea[i];
                thb += sdn[i] * eb[i];
                const int64_t row = (int64_t)ec.sp_slot[i] * S + s0;
                own[row * e.ld] = tha;
                if (lo) lo[row * e.peer_lo_ld] = tha;
                if (hi) hi[row * e.peer_hi_ld] = tha;
                if (s0 + 1 < S) {
                    own[(row + 1) * e.ld] = thb;
                    if (lo) lo[(row + 1) * e.peer_lo_ld] = thb;
                    if (hi) hi[(row + 1) * e.peer_hi_ld] = thb;
                }
            }
        }
#if defined(__CUDA_ARCH__)
        if (lo || hi) __threadfence_system();        // peer stores ordered before this CTA reports completion
#endif
    }

    SVB_HD void store_state(const svbasl_engine &e, int64_t w) const {
        float *s = (e.state_out ? e.state_out : e.state) + w;
        int a = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            s[(int64_t)i * e.ld] = mu[i];
            s[(int64_t)(N + i) * e.ld] = lv[i];
        }
#pragma unroll
        for (int k = 0; k < NL; ++k) s[(int64_t)(2 * N + k) * e.ld] = od[k];
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (e.prior_type[i] == SVBASL_PRIOR_ARD) s[(int64_t)(2 * N + NL + a++) * e.ld] = lphi[i];
    }
};

}  // namespace svb
