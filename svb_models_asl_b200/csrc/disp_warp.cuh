// disp_warp.cuh - fused SVB iteration for aslrest_disp with ONE WARP PER VOXEL.
//
// The dispersion model (model_disp.h; /root/reference/svb_models_asl/aslrest_disp.py:57-110,133-171) costs ~12,000
// instructions per (voxel, sample), almost all of it incomplete-gamma evaluations on the convolution grid, and the work
// of a (voxel, sample) depends on its sampled arrival time (10-45 grid steps), on s (1-8 quadrature pieces per step)
// and on where each grid point sits (series | quadrature | pre-/post-bolus).  With one voxel per thread a warp ran with
// 9 of 32 lanes active (profiles/r1_notes.md section 6).  Here the S x NT (sample, grid point) evaluations of ONE voxel
// are flattened into dense work lists that the 32 lanes go through together, phase by phase:
//
//   0  lane = sample      draws, theta = mu + L eps, transforms, lgamma / digamma of the shape, base value P(a, 2),
//                         the sample's latent-loss terms                                   -> per-sample table (smem)
//   1  lane = (s, k) with x <= 2        fixed 14-term series for P(a,x), J(a,x)            -> P1, J1 (smem)
//   2  lane = (s, k), all grid points   density; 4-point Gauss-Legendre increment of P, J from the previous grid point
//                                       (same piece count for all points of a sample); segmented warp scan = running
//                                       P, J along each sample's grid                      -> P1, J1, D1 (smem)
//   3  lane = (s, k)      AIF and its derivatives (the post-bolus argument s(t - delt - tau) is the grid point tau/h
//                         steps earlier: conv_dt divides tau for every shipped configuration - otherwise the one-thread-
//                         per-voxel kernel is used), tissue recurrence C_k = rho C_{k-1} + dt AIF_k as a segmented warp
//                         scan with weights rho^d, for the value and its three derivatives -> C (smem)
//   4  lane = (s, b)      interpolation at the time points, arterial AIF (its two incomplete-gamma values start from
//                         the nearest grid value below: one short quadrature each), residual, contributions to the
//                         gradient sums (each lane keeps partial sums)
//   5  butterfly reduction of the partial sums; the closing algebra (VoxelStep::finish) and Adam run replicated on the
//      32 lanes, which all store the same values to the same addresses.
//
// The arithmetic of every piece is the one of model_disp.h (same series, same quadrature nodes, same recurrence), so the
// parity tests of the thread-per-voxel kernel apply unchanged (tests/test_kernel_parity.py: disp cases, cuda backend,
// plus a direct comparison of the two kernels).
#pragma once
#include <cstdlib>
#include "kernels.cuh"
#include "model_disp.h"

namespace svb {

#ifndef SVB_DW_PHASE_BARRIER
#define SVB_DW_PHASE_BARRIER 1
#endif
constexpr int kDwTuneDefault = 8;         // CTA-wide barriers (bits 1, 2, 4, 8 = after phases 1, 2, 3, before 5): measured best
constexpr int kDwWarpsDefault = 8;        // voxels (warps) per CTA; the warps of a CTA walk the phases together
constexpr int kDwNtMax = 64;              // convolution grid points supported
constexpr int kDwBMax = 32;               // time points per batch supported (lane = time point)
constexpr int kDwSMax = 32;               // samples supported (lane = sample)

// increment of P and J over [from, x] by n pieces of 4-point Gauss-Legendre (model_disp.h: gamma_run_eval)
__device__ __forceinline__ void gamma_quad(const GammaConst &g, float from, float x, int n, float &dP, float &dJ) {
    const float am1 = g.a - 1.0f;
    const float hw = 0.5f * (x - from) / (float)n;
    const float xi[2] = {0.3399810435848563f, 0.8611363115940526f};
    const float wt[2] = {0.6521451548625461f, 0.3478548451374538f};
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const float mid = from + hw * (float)(2 * k + 1);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float o = hw * xi[j];
            const float ta = mid - o, tb = mid + o;
            const float la = flog(ta), lb = flog(tb);
            const float ea = fexp(am1 * la - ta - g.lg_a), eb = fexp(am1 * lb - tb - g.lg_a);
            s0 += wt[j] * (ea + eb);
            s1 += wt[j] * (ea * la + eb * lb);
        }
    }
    dP = hw * s0;
    dJ = hw * s1;
}

// rarely taken paths, kept out of line: the kernel's instruction footprint decides how often its warps - each in a
// different phase - miss the instruction cache (profiles/r2_notes.md)
static __device__ __noinline__ void igammac_d_slow(const GammaConst &g, float x, float lnx, float &Q, float &dQa, float &dQx) {
    igammac_d(g, x, lnx, Q, dQa, dQx);
}
static __device__ __noinline__ void gamma_series14_slow(const GammaConst &g, float x, float lnx, float &P, float &J) {
    gamma_series14(g, x, lnx, P, J);
}

template <int N, int P>
struct DwRow {                            // per-sample table row (shared memory)
    float x[P > 0 ? P : 1], dx[P > 0 ? P : 1], eps[N], th[N];
    float s, ln_s, sp_live, kct, kcb, delt, deltb, u0, P2, J2, wres, fb, pvf;
    GammaConst g;
    int i0, npts, nsmall, nq, absmode;
};

template <class M>
struct DwLayout {
    static constexpr int N = M::P + 1;
    typedef DwRow<N, M::P> Row;
    // floats per warp: rows, P1/J1/D1 [S][NT], tissue curve at the grid points the time points need [4][S][2B],
    // time-point tables, item offsets, item map (uint16 per (sample, grid point))
    static constexpr int NV = 2 * N + N * (N + 1) / 2 + 1;          // partial sums per lane: a_mu, a_hyp, a_L, cost
    static constexpr int kRowsMax = 80;                              // state rows staged per warp (n_state <= 74)
    static __host__ __device__ size_t grid_floats(int S, int nt) {   // P1/J1/D1, later the reduction tile [NV][33] + totals
        const size_t a = (size_t)3 * S * nt, b = (size_t)NV * 33 + NV;
        return a > b ? a : b;
    }
    static __host__ __device__ size_t floats(int S, int nt, int nb) {
        return (sizeof(Row) * (size_t)S + 3) / 4 + grid_floats(S, nt) + (size_t)8 * S * nb + 4 * kDwBMax + 2 * kDwSMax + 8 +
               (size_t)S * nt + 4 * kRowsMax;
    }
};

// End of a phase.  The shared-memory tables are private to a warp, so __syncwarp() is all correctness needs; the CTA-wide
// barrier keeps the warps of a CTA in the same phase, i.e. in the same stretch of the kernel's ~5,000 instructions: with
// 16 warps per SM each in a phase of its own the instruction cache missed 29 % of its requests and `no_instruction`
// was the top stall (profiles/r2_notes.md).
__device__ __forceinline__ void dw_phase_sync() {
#if SVB_DW_PHASE_BARRIER
    __syncthreads();
#else
    __syncwarp();
#endif
}

// the same, the CTA-wide form selectable at run time (a.tune, identical for every thread): which phase boundaries are
// worth a CTA barrier is a measurement question (waiting for the CTA's slowest warp against instruction-cache locality)
__device__ __forceinline__ void dw_phase_sync_opt(int cta_wide) {
    if (cta_wide) __syncthreads();
    else __syncwarp();
}

__device__ __forceinline__ int warp_incl_scan_int(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// sample that owns flattened item `idx`, given the exclusive offsets off[0..S] in shared memory
__device__ __forceinline__ int dw_find(const int *off, int S, int idx) {
    int s = 0;
    for (int q = 1; q < S; ++q) s += (idx >= off[q]) ? 1 : 0;
    return s;
}

// where a warp's tables live inside its slice of the dynamic shared memory
template <class M>
struct DwCarve {
    typedef DwRow<M::P + 1, M::P> Row;
    Row *rows;
    float *P1, *J1, *D1, *CC, *sm_t, *sm_y, *sm_fr, *st_sm, *m_sm, *v_sm, *g_sm, *scal;
    int *sm_lo, *off_a, *off_b;
    unsigned short *imap, *imap_a;
    __device__ __forceinline__ DwCarve(float *smem, int wv, int S, int nt, int nb) {
        float *base = smem + (size_t)wv * DwLayout<M>::floats(S, nt, nb);
        rows = reinterpret_cast<Row *>(base);
        P1 = base + (sizeof(Row) * (size_t)S + 3) / 4;
        J1 = P1 + (size_t)S * nt;
        D1 = J1 + (size_t)S * nt;
        CC = P1 + DwLayout<M>::grid_floats(S, nt);                                  // CC [4][S][2 nb]
        sm_t = CC + (size_t)8 * S * nb;
        sm_y = sm_t + kDwBMax;
        sm_fr = sm_y + kDwBMax;
        sm_lo = reinterpret_cast<int *>(sm_fr + kDwBMax);
        off_a = sm_lo + kDwBMax;
        off_b = off_a + kDwSMax + 1;
        scal = reinterpret_cast<float *>(off_b + kDwSMax + 1);                      // 6 per-voxel scalars (last, tmax, pv)
        imap = reinterpret_cast<unsigned short *>(off_b + kDwSMax + 1 + 6);         // item -> (s << 6) | k
        imap_a = imap + (size_t)S * nt;                                             // the same for the x <= 2 list
        st_sm = reinterpret_cast<float *>(imap_a + (size_t)S * nt);                 // staged state / m / v / grad rows
        m_sm = st_sm + DwLayout<M>::kRowsMax;
        v_sm = m_sm + DwLayout<M>::kRowsMax;
        g_sm = v_sm + DwLayout<M>::kRowsMax;
    }
};

template <class M, int FL, int kDwWarps>
__global__ void __launch_bounds__(32 * kDwWarps, 16 / kDwWarps) disp_warp_kernel(const __grid_constant__ StepArgs a) {
    extern __shared__ float dw_smem[];
    __shared__ float red[kDwWarps];
    typedef VoxelStep<M, 0, FL> VS;
    constexpr int N = VS::N, P = M::P, NT = VS::NT;
    typedef DwRow<N, P> Row;
    constexpr bool LEAN = FL != 0;
    const svbasl_engine &e = a.e;
    const DevModel &md = a.md;
    const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const int64_t local = (int64_t)blockIdx.x * kDwWarps + wip;
    const bool live = local < e.n_vox;
    const int64_t w = e.w_begin + (live ? local : 0);
    const bool update = LEAN || a.update;
    const int S = e.n_samples, nb = e.n_batch, nt = md.conv_nt;
    const float h = md.conv_h, inv_h = md.conv_inv_h, rho = md.conv_rho, tau = md.tau;
    const int mshift = (int)(tau * inv_h + 0.5f);              // tau / h, an integer (checked by the launcher)

    // ---- warp-private shared memory ----
    const DwCarve<M> cv(dw_smem, wip, S, nt, nb);
    Row *rows = cv.rows;
    float *P1 = cv.P1, *J1 = cv.J1, *D1 = cv.D1, *CC = cv.CC, *sm_t = cv.sm_t, *sm_y = cv.sm_y, *sm_fr = cv.sm_fr;
    int *sm_lo = cv.sm_lo, *off_a = cv.off_a, *off_b = cv.off_b;
    unsigned short *imap = cv.imap, *imap_a = cv.imap_a;
    float *st_sm = cv.st_sm, *m_sm = cv.m_sm, *v_sm = cv.v_sm, *g_sm = cv.g_sm;
    const int nslot = 2 * nb;
    const int n_state = a.n_state;
    // the voxel's state and Adam moments: one row per lane, all in flight at once (the closing algebra reads them from
    // shared memory; row by row from global memory it was a chain of 2 x 36 dependent round trips)
    for (int r = lane; r < n_state; r += 32) {
        st_sm[r] = e.state[(int64_t)r * e.ld + w];
        if (update) {
            m_sm[r] = a.ad.m[(int64_t)r * e.ld + w];
            v_sm[r] = a.ad.v[(int64_t)r * e.ld + w];
        }
    }
    __syncwarp();

    const EngineConst &ec = a.ec;
    const int64_t step = e.step_dev ? (int64_t)*e.step_dev : a.step;
    const int row0 = (update && a.ad.n_batches > 1) ? (int)(step % a.ad.n_batches) : e.t_row0;
    const bool numeric = LEAN || (e.latent == SVBASL_LATENT_NUMERIC);
    const uint32_t key = rng_key(e.seed, step);
    const float pv = md.pvgm ? md.pvgm[w] : md.pvgm_s;

    // ---- time points of the batch: lane = time point ----
    int last = 0;
    float tmax = 0.0f;
    unsigned long long need = 0ull;
    {
        const int64_t stride = (int64_t)e.t_row_stride * e.ld;
        float tv = 0.0f, yv = 0.0f;
        int lo = 0;
        if (lane < nb) {
            yv = e.data[(int64_t)row0 * e.ld + w + lane * stride];
            tv = e.tpts ? e.tpts[(int64_t)row0 * e.ld + w + lane * stride]
                        : e.ti[row0 + lane * e.t_row_stride] + ((e.zoff && !e.tpts) ? e.zoff[w] : 0.0f);
            const float pos = fmin2(fmax2(tv * inv_h, 0.0f), (float)(nt - 1));
            lo = (int)pos;
            lo = lo > nt - 2 ? nt - 2 : lo;
            sm_t[lane] = tv;
            sm_y[lane] = yv;
            sm_lo[lane] = lo;
            sm_fr[lane] = pos - (float)lo;
        }
        last = __reduce_max_sync(FULL, lane < nb ? lo + 1 : 0);
        tmax = __uint_as_float(__reduce_max_sync(FULL, lane < nb ? __float_as_uint(fmax2(tv, 0.0f)) : 0u));
        // grid points whose tissue-curve value some time point interpolates from: bits lo_b and lo_b + 1
        const unsigned m0 = lane < nb ? ((lo < 32 ? 1u << lo : 0u) | (lo + 1 < 32 ? 1u << (lo + 1) : 0u)) : 0u;
        const unsigned m1 = lane < nb ? ((lo >= 32 ? 1u << (lo - 32) : 0u) | (lo + 1 >= 32 ? 1u << (lo + 1 - 32) : 0u)) : 0u;
        need = (unsigned long long)__reduce_or_sync(FULL, m0) | ((unsigned long long)__reduce_or_sync(FULL, m1) << 32);
    }

    // partial sums of this lane: a_mu = sum_s g_s, a_L = sum_s g_s eps_s^T, a_hyp, cost
    float pa_mu[N], pa_L[NT], pa_hyp[N], pcost = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) { pa_mu[i] = 0.0f; pa_hyp[i] = 0.0f; }
#pragma unroll
    for (int k = 0; k < NT; ++k) pa_L[k] = 0.0f;

    // ---- phase 0: lane = sample ----
    // (the posterior state and the prior terms are needed here and in the closing algebra only: they are loaded again
    // there rather than held in ~80 registers through the grid phases)
    int my_small = 0, my_npts = 0;
    // draws of all samples: one Philox call = one posterior row of two samples, the N * ceil(S/2) calls spread over the lanes
    if (!LEAN && e.eps) {
        for (int idx = lane; idx < N * S; idx += 32) {
            const int j = idx / S, s = idx - j * S;
            rows[s].eps[j] = e.eps[((int64_t)j * S + s) * e.ld + w];
        }
    } else {
        const int hp = (S + 1) >> 1;
        for (int p = lane; p < N * hp; p += 32) {
            const int j = p / hp, q = p - j * hp;
            float n0, n1;
            normal_pair(key, e.vox_offset + w, stream_pair(j, 2 * q, S), n0, n1);
            rows[2 * q].eps[j] = n0;
            if (2 * q + 1 < S) rows[2 * q + 1].eps[j] = n1;
        }
    }
    if (lane == 0) {
        cv.scal[0] = __int_as_float(last);
        cv.scal[1] = tmax;
        cv.scal[2] = pv;
    }
    __syncthreads();
    // ---- phase 0, packed over the CTA: thread = (voxel of the CTA, sample) ----
    // One (voxel, sample) costs ~2,000 instructions here (theta, transforms, lgamma / digamma of the shape, the base
    // series at x = 2, the sample's stretch of the grid) and a warp has only S samples for its 32 lanes; packed, the
    // kDwWarps x S pairs of the CTA fill kDwWarps*S/32 warps instead of half-emptying all kDwWarps.
    for (int t = threadIdx.x; t < kDwWarps * S; t += 32 * kDwWarps) {
        const int wv = t / S, s = t - wv * S;
        const DwCarve<M> co(dw_smem, wv, S, nt, nb);
        Row &r = co.rows[s];
        const int last_v = __float_as_int(co.scal[0]);
        const float tmax_v = co.scal[1], pv_v = co.scal[2];
        VS vs;
        vs.load_rows(e, co.st_sm, 1);
        float eps[N], th[N];
#pragma unroll
        for (int j = 0; j < N; ++j) eps[j] = r.eps[j];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float v = vs.mu[i] + fexp(0.5f * vs.lv[i]) * eps[i];
#pragma unroll
            for (int j = 0; j < i; ++j) v += vs.od[stri(i, j)] * eps[j];
            th[i] = v;
            r.th[i] = v;
        }
        float x[P > 0 ? P : 1];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int code = M::xf(p);
            float dxp;
            if (code == SVBASL_XF_EXP) { x[p] = fexp(th[p]); dxp = x[p]; }
            else if (code == SVBASL_XF_ABS) { x[p] = fabsf(th[p]); dxp = th[p] < 0.0f ? -1.0f : 1.0f; }
            else { x[p] = th[p]; dxp = 1.0f; }
            r.x[p] = x[p];
            r.dx[p] = dxp;
        }
        r.wres = ec.scale * fexp(-th[N - 1]);
        // dispersion constants (model_disp.h: prep_disp) and the sample's stretch of the grid
        const float sv = M::DISP ? x[M::ix(M::I_S)] : md.s_fixed;
        const float spv = M::DISP ? x[M::ix(M::I_SP)] : md.sp_fixed;
        r.s = sv;
        r.ln_s = flog(sv);
        r.sp_live = spv < 10.0f ? 1.0f : 0.0f;
        r.g = gamma_const(1.0f + fmin2(spv, 10.0f));
        const float delt = M::ATT ? x[M::ix(M::I_DELT)] : md.att;
        const float deltb = (M::I_DELTBLOOD >= 0) ? x[M::ix(M::I_DELTBLOOD)] : md.artt;
        r.delt = delt;
        r.deltb = deltb;
        r.kct = M::CASL ? 2.0f * fexp(-delt * md.inv_t1b) : 0.0f;
        r.kcb = M::CASL ? 2.0f * fexp(-deltb * md.inv_t1b) : 0.0f;
        r.fb = M::ART ? x[M::ix(M::I_FBLOOD)] : 0.0f;
        r.pvf = pv_v * x[M::ix(M::I_FTISS)];
        const float p0 = delt * inv_h;
        int i0 = p0 > (float)nt ? nt : (p0 > 0.0f ? (int)ceilf(p0) : 0);
        if ((float)i0 * h - delt < 0.0f) ++i0;
        else if (i0 > 0 && (float)(i0 - 1) * h - delt >= 0.0f) --i0;
        r.i0 = i0;
        r.u0 = (float)i0 * h - delt;
        int last_s = last_v;
        if (M::ART) {
            // arterial arguments s (t_b - deltb) are looked up on this grid: extend it far enough
            const float ext = (tmax_v - deltb + delt) * inv_h;
            const int ie = ext > (float)nt ? nt : (ext > 0.0f ? (int)ext + 1 : 0);
            last_s = ie > last_s ? ie : last_s;
        }
        last_s = last_s > nt - 1 ? nt - 1 : last_s;
        const int npts = last_s - i0 + 1 > 0 ? last_s - i0 + 1 : 0;
        r.npts = npts;
        const float wmax = sv * h;
        r.absmode = (wmax > 8.0f || !(wmax > 0.0f)) ? 1 : 0;      // wider than 8 (or degenerate): fresh evaluation per point
        int nq = (int)ceilf(wmax);
        r.nq = nq < 1 ? 1 : (nq > 8 ? 8 : nq);
        int nsm = 0;
        while (nsm < npts && sv * (r.u0 + (float)nsm * h) <= 2.0f) ++nsm;
        r.nsmall = nsm;
        gamma_series14(r.g, 2.0f, 0.6931471805599453f, r.P2, r.J2);
    }
    __syncthreads();
    // the sample's latent-loss and noise terms (voxel_step.h: sample loop), lane = sample of this warp's voxel
    if (lane < S) {
        VS vs;
        vs.load_rows(e, st_sm, 1);
        typename VS::Terms tm;
        vs.prior_terms(e, ec, tm);
        const Row &r = rows[lane];
        float eps[N], th[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { eps[j] = r.eps[j]; th[j] = r.th[j]; }
        pcost += 0.5f * ec.t_full * th[N - 1];
        float g[N];
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] = 0.0f;
        g[N - 1] = 0.5f * ec.t_full;
        if (numeric) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float dth = th[i] - tm.pm[i];
                g[i] += dth * tm.lw_pinv[i];
                pa_hyp[i] += dth * dth;                  // finish() applies lw / v
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            pa_mu[i] += g[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) pa_L[tri(i, j)] += g[i] * eps[j];
        }
        my_small = r.nsmall;
        my_npts = r.npts;
    }
    // exclusive offsets of the two item lists
    {
        const int ia = warp_incl_scan_int(my_small, lane), ib = warp_incl_scan_int(my_npts, lane);
        if (lane < S) { off_a[lane + 1] = ia; off_b[lane + 1] = ib; }
        if (lane == 0) { off_a[0] = 0; off_b[0] = 0; }
    }
    dw_phase_sync();
    const int total_a = off_a[S], total_b = off_b[S];
    if (lane < S) {                                            // every sample lists its own items
        unsigned short *ma = imap_a + off_a[lane], *mb = imap + off_b[lane];
        for (int k = 0; k < my_small; ++k) ma[k] = (unsigned short)((lane << 6) | k);
        for (int k = 0; k < my_npts; ++k) mb[k] = (unsigned short)((lane << 6) | k);
    }
    __syncwarp();

    // ---- phase 1: series for the grid points with x <= 2 ----
    for (int idx = lane; idx < total_a; idx += 32) {
        const int s = imap_a[idx] >> 6, k = imap_a[idx] & 63;
        const Row &r = rows[s];
        const float u = r.u0 + (float)k * h, x = r.s * u;
        float Pv = 0.0f, Jv = 0.0f;
        if (x > 0.0f) gamma_series14(r.g, x, r.ln_s + flog(u), Pv, Jv);
        P1[s * nt + k] = Pv;
        J1[s * nt + k] = Jv;
    }
    dw_phase_sync_opt(a.tune & 1);

    // ---- phase 2: density, quadrature increments, running P and J along each sample's grid ----
    {
        float carryP = 0.0f, carryJ = 0.0f;
        for (int r0 = 0; r0 < total_b; r0 += 32) {
            const int idx = r0 + lane;
            const bool valid = idx < total_b;
            const int sk = valid ? (int)imap[idx] : 0;
            const int s = sk >> 6, k = sk & 63;
            const Row &r = rows[s];
            const float u = r.u0 + (float)k * h, x = r.s * u;
            const float lnx = r.ln_s + flog(fmax2(u, 1e-30f));
            float vP = 0.0f, vJ = 0.0f;
            bool inc = false;
            if (valid) {
                D1[s * nt + k] = x > 0.0f ? fexp((r.g.a - 1.0f) * lnx - x - r.g.lg_a) : 0.0f;
                if (k >= r.nsmall) {
                    if (r.absmode) {
                        float Q, dQa, dQx;
                        igammac_d_slow(r.g, x, lnx, Q, dQa, dQx);
                        P1[s * nt + k] = 1.0f - Q;
                        J1[s * nt + k] = r.g.psi_a * (1.0f - Q) - dQa;
                    } else {
                        const float xp = k > 0 ? r.s * (r.u0 + (float)(k - 1) * h) : 0.0f;
                        const float from = fmax2(xp, 2.0f);
                        if (x - from <= (float)r.nq) {
                            gamma_quad(r.g, from, x, r.nq, vP, vJ);
                        } else {
                            // first point of a sample whose bolus arrived before t = 0 (delt < 0: u0 > h): too wide
                            // for the sample's piece count - evaluate afresh, expressed as the increment over P(a, 2)
                            float Q, dQa, dQx;
                            igammac_d_slow(r.g, x, lnx, Q, dQa, dQx);
                            vP = (1.0f - Q) - r.P2;
                            vJ = (r.g.psi_a * (1.0f - Q) - dQa) - r.J2;
                        }
                        inc = true;
                    }
                }
            }
            // segmented inclusive scan: items of one sample are consecutive, item k has its segment start k lanes back
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tP = __shfl_up_sync(FULL, vP, d), tJ = __shfl_up_sync(FULL, vJ, d);
                if (lane >= d && k >= d) { vP += tP; vJ += tJ; }
            }
            if (k > lane) { vP += carryP; vJ += carryJ; }         // the segment began in an earlier round
            if (inc) {
                P1[s * nt + k] = r.P2 + vP;
                J1[s * nt + k] = r.J2 + vJ;
            }
            carryP = __shfl_sync(FULL, vP, 31);
            carryJ = __shfl_sync(FULL, vJ, 31);
        }
    }
    dw_phase_sync_opt(a.tune & 2);

    // ---- phase 3: AIF on the grid, tissue recurrence as a weighted segmented scan ----
    {
        float rp[5];                                             // rho^1, rho^2, rho^4, rho^8, rho^16
        rp[0] = rho;
#pragma unroll
        for (int q = 1; q < 5; ++q) rp[q] = rp[q - 1] * rp[q - 1];
        const float rho_l1 = fexp2((float)(lane + 1) * (flog(rho) * 1.4426950408889634f));     // rho^(lane+1)
        float carry[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int r0 = 0; r0 < total_b; r0 += 32) {
            const int idx = r0 + lane;
            const bool valid = idx < total_b;
            const int sk = valid ? (int)imap[idx] : 0;
            const int s = sk >> 6, k = sk & 63;
            const Row &r = rows[s];
            float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (valid) {
                const float u = r.u0 + (float)k * h, ti = (float)(r.i0 + k) * h;
                const float Pa = P1[s * nt + k], Ja = J1[s * nt + k];
                const float q1 = 1.0f - Pa, q1a = r.g.psi_a * Pa - Ja, q1x = -D1[s * nt + k];
                float q2 = 1.0f, q2a = 0.0f, q2x = 0.0f;
                const int k2 = k - mshift;
                if (k2 >= 0) {
                    const float Pb = P1[s * nt + k2], Jb = J1[s * nt + k2];
                    q2 = 1.0f - Pb;
                    q2a = r.g.psi_a * Pb - Jb;
                    q2x = -D1[s * nt + k2];
                } else if (u - tau > 0.0f) {
                    // delt < 0: the bolus ended after t = 0 but began before the grid does; no grid value to reuse
                    const float u2 = u - tau;
                    igammac_d_slow(r.g, r.s * u2, r.ln_s + flog(u2), q2, q2a, q2x);
                }
                const bool dead = (md.flags & SVBASL_F_DISP_ASWRITTEN) && ti > r.delt + tau;   // aslrest_disp.py:108
                if (!dead) {
                    const float kc = M::CASL ? r.kct : 2.0f * fexp(-ti * md.inv_t1b);
                    const float dkc = M::CASL ? -kc * md.inv_t1b : 0.0f;
                    const float bq = q2 - q1, u2 = u - tau;
                    c[0] = md.conv_dt * (kc * bq);
                    c[1] = md.conv_dt * (dkc * bq + kc * (-r.s) * (q2x - q1x));
                    c[2] = md.conv_dt * (kc * (q2x * u2 - q1x * u));
                    c[3] = md.conv_dt * (r.sp_live * kc * (q2a - q1a));
                }
            }
#pragma unroll
            for (int q = 0, d = 1; q < 5; ++q, d <<= 1) {
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    const float t = __shfl_up_sync(FULL, c[z], d);
                    if (lane >= d && k >= d) c[z] += rp[q] * t;
                }
            }
            if (k > lane) {
#pragma unroll
                for (int z = 0; z < 4; ++z) c[z] += rho_l1 * carry[z];
            }
            const int ia = r.i0 + k;                                 // absolute grid index
            if (valid && ia >= 0 && ((need >> ia) & 1ull)) {
                const int slot = __popcll(need & ((1ull << ia) - 1ull));
#pragma unroll
                for (int z = 0; z < 4; ++z) CC[((size_t)z * S + s) * nslot + slot] = c[z];
            }
#pragma unroll
            for (int z = 0; z < 4; ++z) carry[z] = __shfl_sync(FULL, c[z], 31);
        }
    }
    dw_phase_sync_opt(a.tune & 4);

    // ---- phase 4: lane = (sample, time point): tissue interpolation, arterial AIF, residual, gradient sums ----
    for (int idx = lane; idx < S * nb; idx += 32) {
        const int s = idx / nb, b = idx - s * nb;
        const Row &r = rows[s];
        const float tb = sm_t[b];
        float pred, d[P > 0 ? P : 1];
#pragma unroll
        for (int p = 0; p < P; ++p) d[p] = 0.0f;
        {
            const int lo = sm_lo[b], klo = lo - r.i0;
            const float fr = sm_fr[b];
            const int slot0 = __popcll(need & ((1ull << lo) - 1ull));      // lo + 1 is needed too: the next slot
            float Sv[4];
#pragma unroll
            for (int z = 0; z < 4; ++z) {
                const float *cz = CC + ((size_t)z * S + s) * nslot;
                const float c0 = (klo >= 0 && klo < r.npts) ? cz[slot0] : 0.0f;
                const float c1 = (klo + 1 >= 0 && klo + 1 < r.npts) ? cz[slot0 + 1] : 0.0f;
                Sv[z] = c0 + fr * (c1 - c0);
            }
            pred = r.pvf * Sv[0];
            d[M::ix(M::I_FTISS)] = pv * Sv[0];
            if (M::ATT) d[M::ix(M::I_DELT)] = r.pvf * Sv[1];
            if (M::DISP) {
                d[M::ix(M::I_S)] = r.pvf * Sv[2];
                d[M::ix(M::I_SP)] = r.pvf * Sv[3];
            }
        }
        if (M::ART) {
            const float v = tb - r.deltb;
            const bool post = tb > r.deltb + tau;
            if (v >= 0.0f && !(post && (md.flags & SVBASL_F_DISP_ASWRITTEN))) {
                float q[2] = {1.0f, 1.0f}, qa[2] = {0.0f, 0.0f}, qx[2] = {0.0f, 0.0f};
                const float vv[2] = {v, v - tau};
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    if (z == 1 && !post) break;
                    const float uu = vv[z], x = r.s * uu;
                    const float lnx = r.ln_s + flog(fmax2(uu, 1e-30f));
                    if (!(x > 0.0f)) continue;                   // Q(a, 0) = 1, no gradient (a > 1)
                    qx[z] = -fexp((r.g.a - 1.0f) * lnx - x - r.g.lg_a);
                    const float kf = floorf((uu - r.u0) * inv_h);
                    const int kk = kf < -1.0f ? -1 : (kf > (float)nt ? nt : (int)kf);
                    float Pv, Jv;
                    if (r.absmode || kk >= r.npts) {
                        float Q, dQa, dQx;
                        igammac_d_slow(r.g, x, lnx, Q, dQa, dQx);
                        Pv = 1.0f - Q;
                        Jv = r.g.psi_a * Pv - dQa;
                    } else {
                        const float xb = kk >= 0 ? r.s * (r.u0 + (float)kk * h) : -1.0f;
                        float from, Pb, Jb;
                        bool quad = true;
                        if (kk >= 0 && xb >= 2.0f) { from = xb; Pb = P1[s * nt + kk]; Jb = J1[s * nt + kk]; }
                        else if (x <= 2.0f) { gamma_series14_slow(r.g, x, lnx, Pb, Jb); from = x; quad = false; }
                        else { from = 2.0f; Pb = r.P2; Jb = r.J2; }
                        float dP = 0.0f, dJ = 0.0f;
                        if (quad && x - from > (float)r.nq) {           // wide gap below the first grid point (delt < 0)
                            float Q, dQa, dQx;
                            igammac_d_slow(r.g, x, lnx, Q, dQa, dQx);
                            Pb = 1.0f - Q;
                            Jb = r.g.psi_a * Pb - dQa;
                            quad = false;
                        }
                        if (quad) gamma_quad(r.g, from, x, r.nq, dP, dJ);
                        Pv = Pb + dP;
                        Jv = Jb + dJ;
                    }
                    q[z] = 1.0f - Pv;
                    qa[z] = r.g.psi_a * Pv - Jv;
                }
                const float kc = M::CASL ? r.kcb : 2.0f * fexp(-tb * md.inv_t1b);
                const float dkc = M::CASL ? -kc * md.inv_t1b : 0.0f;
                const float bq = q[1] - q[0];
                const float A = kc * bq;
                const float dAd = dkc * bq + kc * (-r.s) * (qx[1] - qx[0]);
                const float dAs = kc * (qx[1] * vv[1] - qx[0] * vv[0]);
                const float dAsp = r.sp_live * kc * (qa[1] - qa[0]);
                pred += r.fb * A;
                d[M::ix(M::I_FBLOOD)] = A;
                if (M::I_DELTBLOOD >= 0) d[M::ix(M::I_DELTBLOOD)] = r.fb * dAd;
                if (M::DISP) {
                    d[M::ix(M::I_S)] += r.fb * dAs;
                    d[M::ix(M::I_SP)] += r.fb * dAsp;
                }
            }
        }
        // residual and this element's share of the sums (voxel_step.h: g_p = (T/B)/nu sum_b r_b dpred_b/dx_p T'(theta_p))
        const float res = pred - sm_y[b];
        const float wr = r.wres * res;
        pcost += 0.5f * wr * res;
        float g[N];
#pragma unroll
        for (int p = 0; p < P; ++p) g[p] = wr * d[p] * r.dx[p];
        g[N - 1] = -0.5f * wr * res;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            pa_mu[i] += g[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) pa_L[tri(i, j)] += g[i] * r.eps[j];
        }
    }

    // ---- phase 5: sums over the lanes through a shared-memory tile, closing algebra, update ----
    {
        constexpr int NV = DwLayout<M>::NV;
        float *tile = P1, *tots = P1 + NV * 33;                 // P1 / J1 / D1 are no longer needed
        dw_phase_sync_opt(a.tune & 8);
#pragma unroll
        for (int i = 0; i < N; ++i) { tile[i * 33 + lane] = pa_mu[i]; tile[(N + i) * 33 + lane] = pa_hyp[i]; }
#pragma unroll
        for (int k = 0; k < NT; ++k) tile[(2 * N + k) * 33 + lane] = pa_L[k];
        tile[(NV - 1) * 33 + lane] = pcost;
        __syncwarp();
        for (int q = lane; q < NV; q += 32) {
            const float *row = tile + q * 33;
            float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;   // same order on every launch: bit-reproducible
#pragma unroll
            for (int l = 0; l < 32; l += 4) { t0 += row[l]; t1 += row[l + 1]; t2 += row[l + 2]; t3 += row[l + 3]; }
            tots[q] = (t0 + t1) + (t2 + t3);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < N; ++i) { pa_mu[i] = tots[i]; pa_hyp[i] = tots[N + i]; }
#pragma unroll
        for (int k = 0; k < NT; ++k) pa_L[k] = tots[2 * N + k];
        pcost = tots[NV - 1];
    }
    VS vs;
    vs.load_rows(e, st_sm, 1);
    typename VS::Terms tm;
    vs.prior_terms(e, ec, tm);
    float cost = vs.finish(e, ec, tm, pa_mu, pa_L, pa_hyp, pcost, numeric);
    int skipped = 0;
    if (live) {
        if (!LEAN && a.cost) a.cost[w] = cost;
        if (!LEAN && a.grad) vs.store_grads(e, a.grad, w);
        if (update) {
            // TF-Adam, one state row per lane (the gradient rows go through shared memory in state order)
            const bool ok = vs.grads_finite() && cost == cost;
            if (lane == 0) vs.store_grad_rows(e, g_sm, 1);
            __syncwarp();
            float *so = (e.state_out ? e.state_out : e.state) + w;
            const float lr_t = a.ad.lr_t[step];
            for (int r = lane; r < n_state; r += 32) {
                if (ok) {
                    float mm = m_sm[r], vv = v_sm[r];
                    so[(int64_t)r * e.ld] = VS::adam1(a.ad, lr_t, st_sm[r], g_sm[r], mm, vv);
                    a.ad.m[(int64_t)r * e.ld + w] = mm;
                    a.ad.v[(int64_t)r * e.ld + w] = vv;
                } else if (e.state_out && e.state_out != e.state) {
                    so[(int64_t)r * e.ld] = st_sm[r];
                }
            }
            if (!ok) {
                skipped = lane == 0 ? 1 : 0;
                cost = 0.0f;
            }
        }
    } else {
        cost = 0.0f;
    }
    if (a.cost_sum) {
        if (lane == 0) red[wip] = cost;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
#pragma unroll
            for (int i = 0; i < kDwWarps; ++i) t += red[i];
            atomicAdd(a.cost_sum + ((e.step_dev && !e.cost_sum_scalar) ? step : 0), (double)t);
        }
    }
    if (a.nan_count && skipped) atomicAdd((unsigned long long *)a.nan_count, 1ull);
}

// Does this call fit the warp-per-voxel kernel?  (tissue component, voxel-wise priors, sizes within one warp's lanes,
// tau a whole number of grid steps, one iteration per launch)
inline bool disp_warp_eligible(const StepArgs &a, bool tissue) {
    if (!tissue || a.e.n_samples > kDwSMax || a.e.n_batch > kDwBMax || a.md.conv_nt > kDwNtMax || a.md.conv_nt < 2) return false;
    if (a.update && a.ad.n_iters > 1 && a.e.step_dev) return false;   // fused iterations are launched one by one from the host
    for (int i = 0; i < a.e.n_par && i < SVBASL_MAX_PAR; ++i)
        if (a.e.prior_type[i] == SVBASL_PRIOR_MRF) return false;
    const float m = a.md.tau * a.md.conv_inv_h;
    const float r = floorf(m + 0.5f);
    if (!(r >= 1.0f) || fabsf(m - r) > 1e-4f * r) return false;
    if (fabsf(a.md.conv_dt - a.md.conv_h) > 1e-6f * a.md.conv_h) return false;     // linspace step == conv_dt
    return true;
}

template <class M, int FL, int kDwWarps>
int launch_step_disp_warp(const StepArgs &a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.e.n_vox + kDwWarps - 1) / kDwWarps);
    if (grid == 0) return 0;
    const size_t smem = sizeof(float) * kDwWarps * DwLayout<M>::floats(a.e.n_samples, a.md.conv_nt, a.e.n_batch);
    if (smem > 48 * 1024) {
        static std::atomic<int> granted[kMaxDevices];
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = -1;
        int have = dev >= 0 ? granted[dev].load(std::memory_order_acquire) : 0;
        if (have == 0) {
            int optin = 0;
            cudaFuncAttributes fa;
            cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev < 0 ? 0 : dev);
            if (err == cudaSuccess) err = cudaFuncGetAttributes(&fa, disp_warp_kernel<M, FL, kDwWarps>);
            if (err == cudaSuccess) {
                have = optin - (int)fa.sharedSizeBytes;
                err = cudaFuncSetAttribute(disp_warp_kernel<M, FL, kDwWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, have);
            }
            if (err != cudaSuccess) {
                set_error("disp_warp_kernel: cannot opt in to large shared memory: %s", cudaGetErrorString(err));
                return SVBASL_E_CUDA;
            }
            if (dev >= 0) granted[dev].store(have, std::memory_order_release);
        }
        if ((size_t)have < smem) return 1;                         // does not fit: caller tries fewer warps / the scalar kernel
    }
    StepArgs b = a;
    const char *tune = getenv("SVBASL_DW_TUNE");                    // measurement switch: bit mask of CTA-wide phase barriers
    b.tune = tune ? atoi(tune) : kDwTuneDefault;
    disp_warp_kernel<M, FL, kDwWarps><<<grid, 32 * kDwWarps, smem, st>>>(b);
    return check_launch("disp_warp_kernel");
}

// launcher stored in the dispatch table for aslrest_disp: warp-per-voxel when the call fits, else thread-per-voxel
template <class M, int FL>
int launch_step_disp(const StepArgs &a, cudaStream_t st) {
    const bool force_scalar = getenv("SVBASL_DISP_SCALAR") != nullptr;              // tests compare the two kernels
    if (!force_scalar && disp_warp_eligible(a, M::TISS)) {
        // 8 warps per CTA (two CTAs per SM); 4 when the per-warp tables of a large S / grid / batch do not fit
        const char *wenv = getenv("SVBASL_DW_WARPS");               // measurement switch
        // svbasl_adam.n_iters fused iterations: the warp kernel keeps nothing on chip between iterations that would
        // pay for fusing them (its state rows are 1 % of its time), so they are n_iters launches of the same call
        const int n_iters = a.update ? a.ad.n_iters : 1;
        StepArgs b = a;
        b.ad.n_iters = 1;
        for (int it = 0; it < n_iters; ++it) {
            b.step = a.step + it;
            b.cost_sum = a.cost_sum ? a.cost_sum + it : nullptr;
            int rc = 1;
            if (wenv && wenv[0] == '1' && wenv[1] == '6') rc = launch_step_disp_warp<M, FL, 16>(b, st);
            if (rc > 0 && !(wenv && wenv[0] == '4')) rc = launch_step_disp_warp<M, FL, kDwWarpsDefault>(b, st);
            if (rc > 0) rc = launch_step_disp_warp<M, FL, 4>(b, st);
            if (rc < 0) return rc;
            if (rc > 0) {
                if (it == 0) break;                                 // does not fit at all: the thread-per-voxel kernel
                return SVBASL_E_CUDA;
            }
            if (it == n_iters - 1) return 0;
            if (b.e.state_out) b.e.state = b.e.state_out;           // the next iteration continues from what this one wrote
        }
    }
    return launch_step<M, 0, FL>(a, st);
}

}  // namespace svb
