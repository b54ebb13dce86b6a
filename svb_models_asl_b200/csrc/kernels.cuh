// kernels.cuh - __global__ wrappers around the per-voxel bodies (voxel_step.h, model_*.h) and their
// launchers.  One thread owns one voxel: every load/store of the SoA arrays is a fully coalesced 128-byte
// row segment per warp, and all per-sample work stays in registers.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <type_traits>
#include "voxel_step.h"

namespace svb {

#ifndef SVB_KBLOCK
#define SVB_KBLOCK 128
#endif
constexpr int kBlock = SVB_KBLOCK;               // voxels (threads) per CTA
constexpr int kMaxDevices = 64;                  // per-device caches of function attributes

struct StepArgs {
    DevModel md;
    svbasl_engine e;
    EngineConst ec;
    svbasl_adam ad;
    int32_t update;          // 0: cost + gradient only (svbasl_elbo_grad); 1: fused Adam update (svbasl_step)
    int32_t n_state;         // rows of state / m / v
    int32_t nb_param;        // spatial parameter whose neighbour samples are staged in shared memory, -1 = none
                             // (chosen by launch_step)
    int64_t step;            // iteration index of the first fused iteration (RNG counter / lr_t index)
    float *cost;             // [ld] or NULL
    float *grad;             // [n_state][ld] or NULL
    double *cost_sum;        // [n_iters] or NULL
    long long *nan_count;    // [1] or NULL
    svbasl_hyper hy;         // fused tail of a spatial iteration; hy.done_ctas == NULL: none
    int32_t tune;            // disp_warp_kernel: which phase boundaries are CTA-wide barriers
};

// Mailbox of one rank for the all-reduce of d(cost)/d(log ak) over NVLink peer memory: [2 parities][world] slots;
// slot q of parity p is written by rank q only.
struct MailSlot {
    double val[SVBASL_MAX_SPATIAL];
    unsigned long long seq;
};

// All-reduce of ak_grad over the ranks' mailboxes + TF-Adam step on log ak + reset of the accumulators + advance of
// the iteration counter.  Called by ONE CTA (threads r < world each serve one rank) once every step kernel of this
// rank's iteration has finished: as the tail of the step kernel's last CTA (svbasl_step_spatial) or as a launch of
// its own (svbasl_hyper_step_peers).  Every rank stores its partial sums and a sequence number (= step + 1) into ITS
// slot of every rank's mailbox (double-buffered by parity), waits - bounded - until all slots of its own mailbox carry
// that number, and adds them in rank order in double precision: bit-identical on every rank.  The flag is stored
// behind a system-scope fence, i.e. behind this GPU's earlier peer stores (the mirrored halo samples), so the same
// wait is the barrier that makes the neighbours' halo data visible before the next iteration.
__device__ __forceinline__ void hyper_tail(const svbasl_hyper &hy, double *ak_grad, float grad_scale, double (*part)[SVBASL_MAX_SPATIAL]) {
    const int r = threadIdx.x;
    const int n = hy.n_spatial, world = hy.world;
    const long long step = *hy.step_dev;
    if (world > 1) {
        const unsigned long long seq = (unsigned long long)step + 1ull;
        const int par = (int)(seq & 1ull);
        if (r < world) {
            MailSlot *dst = (MailSlot *)hy.mailboxes[r] + (size_t)par * world + hy.rank;
            for (int k = 0; k < n; ++k) ((volatile double *)dst->val)[k] = ak_grad[k];
            __threadfence_system();                   // values (and this GPU's earlier peer stores) before the flag
            *(volatile unsigned long long *)&dst->seq = seq;
            const MailSlot *src = (const MailSlot *)hy.mailboxes[hy.rank] + (size_t)par * world + r;
            const volatile unsigned long long *flag = &src->seq;
            if (*(volatile int *)hy.status == 0) {
                const long long t0 = clock64();
                while (*flag != seq) {
                    if (clock64() - t0 > 20000000000ll) {   // ~10 s at 1.9 GHz
                        atomicExch(hy.status, 1);
                        break;
                    }
                }
            }
            __threadfence_system();
            for (int k = 0; k < n; ++k) part[r][k] = ((const volatile double *)src->val)[k];
        }
    } else if (r == 0) {
        for (int k = 0; k < n; ++k) part[0][k] = ak_grad[k];
    }
    __syncthreads();
    if (r < n) {
        double sum = 0.0;
        for (int q = 0; q < world; ++q) sum += part[q][r];
        const float g = (float)(sum * (double)grad_scale);
        const float mm = hy.beta1 * hy.m[r] + (1.0f - hy.beta1) * g;
        const float vv = hy.beta2 * hy.v[r] + (1.0f - hy.beta2) * g * g;
        hy.m[r] = mm;
        hy.v[r] = vv;
        hy.log_ak[r] -= hy.lr_t[step] * mm / (sqrtf(vv) + hy.epsilon);
        ak_grad[r] = 0.0;
    }
    __syncthreads();
    if (r == 0) *hy.step_dev = step + 1;
}

struct EvalArgs {
    DevModel md;
    const float *params;     // [P][n_rows]
    const float *tpts;       // [n_t_rows][B]
    float *out;              // [n_rows][B]
    int64_t n_rows, n_t_rows;
    int32_t n_samples, n_batch;
};

struct FitArgs {
    DevModel md;
    svbasl_engine e;
    float *out;              // [T][ld]
};

typedef int (*step_launcher_t)(const StepArgs &, cudaStream_t);
typedef int (*eval_launcher_t)(const EvalArgs &, cudaStream_t);
typedef int (*fit_launcher_t)(const FitArgs &, cudaStream_t);

struct KernelEntry {
    int32_t kind;
    uint32_t flags;          // canonical SVBASL_F_* set
    int32_t nbt;             // compile-time batch size, 0 = any
    int32_t lean;            // flavour: 0 generic; 1 production (update, Philox, numeric latent loss, no spatial
                             // prior, no per-voxel outputs); 2 production with the spatial prior
    int32_t n_params;
    step_launcher_t step;
    eval_launcher_t eval;    // only on the nbt == 0, mrfmask == 0 entry
    fit_launcher_t fit;
};

// models whose run() is a CTA-cooperative routine (model_nn_tc.cuh) bracket the kernel with cta_begin / cta_end
template <class M, class = void>
struct is_cta_coop : std::false_type {};
template <class M>
struct is_cta_coop<M, std::void_t<decltype(M::kCtaCoop)>> : std::bool_constant<M::kCtaCoop> {};

constexpr int kMaxDeferredCosts = 16;
#ifndef SVB_WRITE_BACK_AFTER
#define SVB_WRITE_BACK_AFTER 1
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum -> one double atomic per block
__device__ __forceinline__ void block_accumulate(float v, double *dst, float *smem) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < kBlock / 32; ++i) t += smem[i];
        atomicAdd(dst, (double)t);
    }
    __syncthreads();
}

// 4-byte asynchronous global->shared copy (LDGSTS): the data lands in shared memory without passing through
// a register, so a load issued before the sample loop costs nothing while the loop runs.
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const float *gsrc) {     // smem_dst: shared-window address
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Fused iteration.  Dynamic shared memory: [2][n_state][kBlock] floats when updating (the Adam moments of this
// CTA's voxels, prefetched asynchronously at kernel start and consumed after the sample loop), followed - when
// a.nb_param >= 0 - by [S][6][kBlock] floats: the six neighbours' samples of spatial parameter nb_param, gathered
// asynchronously so that the sample loop reads them from shared memory.
// Resident CTAs per SM the register allocation is tuned for: 4 (<= 128 registers) for the common layouts,
// fewer for the wide posteriors (P' >= 6: the Cholesky factor and its gradient alone are P'(P'+1) registers).
template <class M>
constexpr int min_blocks() {
#ifdef SVB_MIN_BLOCKS                     // tuning builds (scratch/build_variant.sh)
    return SVB_MIN_BLOCKS;
#endif
    return (128 / kBlock) * (M::kRegHeavy ? 3 : (M::P + 1 <= 5 ? 4 : (M::P + 1 <= 7 ? 3 : 2)));
}

// FL != 0: production flavours - fused update, no per-voxel cost / gradient outputs (see VoxelStep)
template <class M, int NBT, int FL>
#ifdef SVB_MAXNREG                        // tuning builds: an explicit register cap instead of the resident-CTA hint
__global__ void __maxnreg__(SVB_MAXNREG) step_kernel(const __grid_constant__ StepArgs a) {
#else
__global__ void __launch_bounds__(kBlock, min_blocks<M>()) step_kernel(const __grid_constant__ StepArgs a) {
#endif
    extern __shared__ float mv_tile[];
    __shared__ float red[kBlock / 32];
    __shared__ float red_it[kMaxDeferredCosts][kBlock / 32];
    typedef VoxelStep<M, NBT, FL> VS;
    constexpr bool LEAN = FL != 0;
    constexpr bool SPATIAL = FL != 1;
    // Production flavour without the spatial prior (the one that fuses iterations): every iteration updates the moments
    // in their shared-memory tile and the state in registers, and ONE copy loop after the iteration loop writes both
    // back - the iteration loop then holds a single instance of the Adam code (instruction-cache footprint).
    constexpr bool WRITE_BACK_AFTER = SVB_WRITE_BACK_AFTER && FL == 1;
    const bool update = LEAN || a.update;
    if constexpr (is_cta_coop<M>::value) M::cta_begin(a.md);
    const int64_t local = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool live = local < a.e.n_vox;
    int64_t idx = live ? local : 0;
    if (SPATIAL && (a.e.peer_lo || a.e.peer_hi)) {
        // shard-boundary voxels (the first peer_lo_count and the last peer_hi_count owned voxels) are given to the
        // first CTAs, so their stores into the neighbour ranks' halos cross NVLink while the interior computes
        const int64_t nlo = a.e.peer_lo ? a.e.peer_lo_count : 0, nhi = a.e.peer_hi ? a.e.peer_hi_count : 0;
        if (idx >= nlo) idx = idx < nlo + nhi ? a.e.n_vox - nhi + (idx - nlo) : idx - nhi;
    }
    const int64_t w = a.e.w_begin + idx;
    const int n_state = a.n_state;
    float *m_sm = mv_tile + threadIdx.x;
    float *v_sm = mv_tile + (size_t)n_state * kBlock + threadIdx.x;
    // Prologue order: every independent global load is issued before anything waits on one - the six neighbour
    // indices, then the posterior state (vs.load), then the asynchronous copies.  (With the index loads inside
    // the copy loop a warp spent 14 % of its life on six serialised DRAM round trips, profiles/r1_notes.md.)
    int nb_u[6];
    const bool tile = SPATIAL && a.nb_param >= 0;
    if (tile) {
#pragma unroll
        for (int k = 0; k < 6; ++k) nb_u[k] = a.e.neighbours[(int64_t)k * a.e.ld + w];
    }
    VS vs;
    vs.load(a.e, w);
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(mv_tile) + 4u * threadIdx.x;
    constexpr unsigned kRow = 4u * kBlock;                       // bytes per shared-memory row
    if (update) {
        const float *mg = a.ad.m + w, *vg = a.ad.v + w;
        unsigned dm = sm0, dv = sm0 + (unsigned)n_state * kRow;
        for (int k = 0; k < n_state; ++k, dm += kRow, dv += kRow, mg += a.e.ld, vg += a.e.ld) {
            cp_async4(dm, mg);
            cp_async4(dv, vg);
        }
    }
    NbTile nbt = {nullptr, 0, -1};
    if (tile) {
        const size_t off = update ? 2 * (size_t)n_state * kBlock : 0;
        const int S = a.e.n_samples;
        const float *src = a.e.spatial_samples + (int64_t)a.ec.sp_slot[a.nb_param] * S * a.e.ld;
        unsigned dst = sm0 + (unsigned)off * 4u;
        // A missing neighbour (volume edge, masked-out voxel) is replaced by the voxel ITSELF: the buffer holds the
        // voxel's own sample of this iteration too - the very value the sample loop computes for theta (bit for bit for
        // the first parameter, to rounding otherwise) - so its difference term vanishes and neither the copies nor the
        // sample loop test a mask.
#pragma unroll
        for (int k = 0; k < 6; ++k) nb_u[k] = nb_u[k] >= 0 ? nb_u[k] : (int)w;
        for (int s = 0; s < S; ++s, src += a.e.ld, dst += 6u * kRow) {
#pragma unroll
            for (int k = 0; k < 6; ++k) cp_async4(dst + k * kRow, src + nb_u[k]);
        }
        nbt.v = mv_tile + off + threadIdx.x;
        nbt.stride = kBlock;
        nbt.param = a.nb_param;
    }
    const int n_iters = update ? a.ad.n_iters : 1;
    // the per-iteration cost sums of a fused launch leave the CTA together after the last iteration: no CTA barrier
    // inside the iteration loop
    const bool defer_costs = n_iters > 1 && n_iters <= kMaxDeferredCosts;
    int skipped = 0;
    const int64_t step_base = a.e.step_dev ? (int64_t)*a.e.step_dev : a.step;   // device counter under graph replay
    double *cost_sum = a.cost_sum ? a.cost_sum + ((a.e.step_dev && !a.e.cost_sum_scalar) ? step_base : 0) : nullptr;
    // the batch (data, time points, what the model derives from them) is loaded when its row changes: once per launch
    // when all fused iterations see the same batch
    typename M::Vox vox = M::load_vox(a.md, w);
    BatchAcc<VS::P, NBT> acc;
    int row_loaded = -1;
    for (int it = 0; it < n_iters; ++it) {
        const int64_t step = step_base + it;
        const int row0 = (update && a.ad.n_batches > 1) ? (int)(step % a.ad.n_batches) : a.e.t_row0;
        if (row0 != row_loaded) {
            acc.load(a.e, w, row0);
            M::bind_times(a.md, vox, acc);
            row_loaded = row0;
        }
        float cost = vs.elbo_grad_batch(a.md, a.e, a.ec, w, step, vox, acc, nbt);
        if constexpr (is_cta_coop<M>::value) {
            if (M::cta_failed()) cost = nanf("");            // a tensor-core wait expired: no update, counted as skipped
        }
        if (update && it == 0) cp_async_wait_all();          // only this thread reads what it copied: no barrier
        if (live) {
            if (!LEAN && a.cost) a.cost[w] = cost;
            if (!LEAN && a.grad) vs.store_grads(a.e, a.grad, w);
            if (update) {
                if (vs.grads_finite() && cost == cost) {
                    // the moments live in the shared-memory tile they were prefetched into for the whole launch;
                    // the last fused iteration writes them (and the state) back to global memory
                    if (!WRITE_BACK_AFTER && it == n_iters - 1)
                        vs.adam_update(a.e, a.ad, a.ad.lr_t[step], w, true, m_sm, v_sm, kBlock, a.ad.m + w, a.ad.v + w, a.e.ld);
                    else
                        vs.adam_update(a.e, a.ad, a.ad.lr_t[step], w, false, m_sm, v_sm, kBlock, m_sm, v_sm, kBlock);
                    if (SPATIAL && it == n_iters - 1 && a.e.spatial_samples_out) vs.store_next_samples(a.e, a.ec, w, step + 1);
                } else {
                    ++skipped;
                    if (!WRITE_BACK_AFTER && it == n_iters - 1) {
                        vs.store_state(a.e, w);
                        // the next iteration's samples are still due (unchanged state, new draws): the sample
                        // buffers - our own and the neighbour rank's halo columns - ping-pong every iteration
                        if (SPATIAL && a.e.spatial_samples_out) vs.store_next_samples(a.e, a.ec, w, step + 1);
                        if (n_iters > 1) {                     // moments of the earlier fused iterations
                            for (int k = 0; k < n_state; ++k) {
                                a.ad.m[(int64_t)k * a.e.ld + w] = m_sm[k * kBlock];
                                a.ad.v[(int64_t)k * a.e.ld + w] = v_sm[k * kBlock];
                            }
                        }
                    }
                    cost = 0.0f;
                }
            }
        } else {
            cost = 0.0f;
        }
        if (cost_sum) {
            if (defer_costs) {                               // warp sums now, the CTA's sum and its atomic after the loop
                const float ws = warp_sum(cost);
                if ((threadIdx.x & 31) == 0) red_it[it][threadIdx.x >> 5] = ws;
            } else {
                block_accumulate(cost, cost_sum + it, red);
            }
        }
        if (SPATIAL && a.e.ak_grad) {
#pragma unroll
            for (int i = 0; i < VS::N; ++i)
                if (a.e.prior_type[i] == SVBASL_PRIOR_MRF)
                    block_accumulate(live ? vs.ak_out[i] : 0.0f, a.e.ak_grad + a.ec.sp_slot[i], red);
        }
    }
    if (WRITE_BACK_AFTER && update && live) {
        vs.store_state(a.e, w);
        float *mg = a.ad.m + w, *vg = a.ad.v + w;
        for (int k = 0; k < n_state; ++k, mg += a.e.ld, vg += a.e.ld) {
            *mg = m_sm[k * kBlock];
            *vg = v_sm[k * kBlock];
        }
    }
    if (cost_sum && defer_costs) {
        __syncthreads();
        if (threadIdx.x < n_iters) {
            float t = 0.0f;
#pragma unroll
            for (int i = 0; i < kBlock / 32; ++i) t += red_it[threadIdx.x][i];
            atomicAdd(cost_sum + threadIdx.x, (double)t);
        }
    }
    if (a.nan_count && skipped) atomicAdd((unsigned long long *)a.nan_count, (unsigned long long)skipped);
    if constexpr (is_cta_coop<M>::value) M::cta_end();
    if constexpr (SPATIAL) {
        if (a.hy.done_ctas) {
            // Fused tail of the iteration: the last CTA of this launch to get here owns the hyper-parameter step.
            // (threadFenceReduction pattern: every thread's global / peer stores are fenced, the CTA joins, thread 0
            // publishes the CTA's completion; the CTA that observes the full count has all others' results visible.)
            __shared__ int is_last;
            __shared__ double part[SVBASL_MAX_PEERS][SVBASL_MAX_SPATIAL];
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned prev = atomicAdd(a.hy.done_ctas, 1u);
                is_last = prev == gridDim.x - 1;
                if (is_last) *a.hy.done_ctas = 0u;
            }
            __syncthreads();
            if (is_last) {
                __threadfence();
                hyper_tail(a.hy, a.e.ak_grad, a.e.grad_scale, part);
            }
        }
    }
}

// Pre-pass for spatial priors: theta samples of every spatially-regularised parameter for the voxels
// [first, first + count) of the local arrays (all local voxels incl. the halo when priming, the owned voxels when it
// runs right behind a step to prepare the next iteration).  A shard-boundary voxel's samples are ALSO stored into the
// adjacent rank's halo columns through NVLink peer memory (svbasl_engine.peer_*): the halo "exchange" is these
// stores.  Boundary voxels are served by the first CTAs so the stores are under way early.
struct SpatialArgs {
    svbasl_engine e;
    EngineConst ec;
    int64_t first, count;
    int64_t step;            // added to *e.step_dev when the iteration counter lives on the device
    float *out;              // [n_spatial][S][ld]
};

static __global__ void __launch_bounds__(kBlock) spatial_sample_kernel(const __grid_constant__ SpatialArgs a) {
    int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (idx >= a.count) return;
    const bool peers = a.e.peer_lo || a.e.peer_hi;
    if (peers) {
        const int64_t nlo = a.e.peer_lo ? a.e.peer_lo_count : 0, nhi = a.e.peer_hi ? a.e.peer_hi_count : 0;
        if (idx >= nlo) idx = idx < nlo + nhi ? a.count - nhi + (idx - nlo) : idx - nhi;
    }
    const int64_t u = a.first + idx;
    const uint32_t key = rng_key(a.e.seed, (a.e.step_dev ? (int64_t)*a.e.step_dev : 0) + a.step);
    const int S = a.e.n_samples, n = a.e.n_par;
    const int64_t ld = a.e.ld;
    float *lo = (a.e.peer_lo && u >= a.e.peer_lo_first && u < a.e.peer_lo_first + a.e.peer_lo_count)
                    ? a.e.peer_lo + (u + a.e.peer_lo_shift) : nullptr;
    float *hi = (a.e.peer_hi && u >= a.e.peer_hi_first && u < a.e.peer_hi_first + a.e.peer_hi_count)
                    ? a.e.peer_hi + (u + a.e.peer_hi_shift) : nullptr;
    auto put = [&](int slot, int s, float th) {
        const int64_t row = (int64_t)slot * S + s;
        a.out[row * ld + u] = th;
        if (lo) lo[row * a.e.peer_lo_ld] = th;
        if (hi) hi[row * a.e.peer_hi_ld] = th;
    };
    int p_first = 0;
    if (!a.e.eps) {
        // Parameters 0 and 1 (ftiss, delttiss: the usual spatial ones): one Philox call per row serves two samples;
        // the row of L is read once (same arithmetic, term by term, as sample_theta / store_next_samples).
        const int s0 = a.ec.sp_slot[0], s1 = n > 1 ? a.ec.sp_slot[1] : -1;
        p_first = n > 1 ? 2 : 1;
        if (s0 >= 0 || s1 >= 0) {
            const float *st = a.e.state + u;
            const float mu0 = st[0], sd0 = fexp(0.5f * st[(int64_t)n * ld]);
            float mu1 = 0.0f, sd1 = 0.0f, od10 = 0.0f;
            if (s1 >= 0) {
                mu1 = st[ld];
                sd1 = fexp(0.5f * st[(int64_t)(n + 1) * ld]);
                od10 = st[(int64_t)(2 * n + stri(1, 0)) * ld];
            }
            for (int s = 0; s < S; s += 2) {
                float e0a, e0b, e1a = 0.0f, e1b = 0.0f;
                normal_pair(key, a.e.vox_offset + u, stream_pair(0, s, S), e0a, e0b);
                if (s1 >= 0) normal_pair(key, a.e.vox_offset + u, stream_pair(1, s, S), e1a, e1b);
                const bool two = s + 1 < S;
                if (s0 >= 0) {
                    put(s0, s, mu0 + sd0 * e0a);
                    if (two) put(s0, s + 1, mu0 + sd0 * e0b);
                }
                if (s1 >= 0) {
                    put(s1, s, (mu1 + od10 * e0a) + sd1 * e1a);
                    if (two) put(s1, s + 1, (mu1 + od10 * e0b) + sd1 * e1b);
                }
            }
        }
    }
    for (int p = p_first; p < n; ++p) {
        const int slot = a.ec.sp_slot[p];
        if (slot < 0) continue;
        for (int s = 0; s < S; ++s) put(slot, s, sample_theta(a.e, key, u, p, s));
    }
    if (lo || hi) __threadfence_system();
}

// Model.evaluate: one thread per (row, time point) element of the reference's [W,S,B] output
template <class M>
__global__ void __launch_bounds__(256) eval_kernel(const __grid_constant__ EvalArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = a.n_rows * a.n_batch;
    if (idx >= total) return;
    const int64_t row = idx / a.n_batch;
    const int b = (int)(idx - row * a.n_batch);
    const int64_t vox = row / a.n_samples;
    const int64_t rows_per_t = a.n_rows / a.n_t_rows;
    const float t = a.tpts[(row / rows_per_t) * a.n_batch + b];
    float x[M::P > 0 ? M::P : 1];
#pragma unroll
    for (int p = 0; p < M::P; ++p) x[p] = a.params[(int64_t)p * a.n_rows + row];
    typename M::Vox vx = M::load_vox(a.md, vox);
    a.out[idx] = M::predict(a.md, vx, x, t);
}

// Prediction at the posterior mean for every time point (save_model_fit, asl_example.py:39): out [T][ld]
template <class M>
__global__ void __launch_bounds__(kBlock) fit_kernel(const __grid_constant__ FitArgs a) {
    const int64_t local = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (local >= a.e.n_vox) return;
    const int64_t w = a.e.w_begin + local;
    float x[M::P > 0 ? M::P : 1];
#pragma unroll
    for (int p = 0; p < M::P; ++p) {
        float th = a.e.state[(int64_t)p * a.e.ld + w];
        const int code = M::xf(p);
        x[p] = code == SVBASL_XF_EXP ? fexp(th) : (code == SVBASL_XF_ABS ? fabsf(th) : th);
    }
    typename M::Vox vx = M::load_vox(a.md, w);
    const float zoff = (a.e.zoff && !a.e.tpts) ? a.e.zoff[w] : 0.0f;
    for (int r = 0; r < a.e.t_full; ++r) {
        const float t = a.e.tpts ? a.e.tpts[(int64_t)r * a.e.ld + w] : a.e.ti[r] + zoff;
        a.out[(int64_t)r * a.e.ld + w] = M::predict(a.md, vx, x, t);
    }
}

inline int check_launch(const char *what);
void set_error(const char *fmt, ...);

template <class M, int NBT, int FL>
int launch_step(const StepArgs &a0, cudaStream_t st) {
    StepArgs a = a0;
    const unsigned grid = (unsigned)((a.e.n_vox + kBlock - 1) / kBlock);
    if (grid == 0) return 0;
    size_t smem = a.update ? sizeof(float) * 2 * (size_t)a.n_state * kBlock : 0;
    a.nb_param = -1;
    if (FL != 1) {
        // neighbour tile of the first spatial parameter, if the planned number of resident CTAs still fits
        const size_t tile = sizeof(float) * 6 * (size_t)a.e.n_samples * kBlock;
        for (int i = 0; i < a.e.n_par && i < SVBASL_MAX_PAR && a.nb_param < 0; ++i)
            if (a.e.prior_type[i] == SVBASL_PRIOR_MRF && (smem + tile + 1024) * min_blocks<M>() <= 227 * 1024) a.nb_param = i;
        if (a.nb_param >= 0) smem += tile;
    }
    // The opt-in above 48 KB is a per-device (per-context) function attribute.  It is raised ONCE per
    // (instantiation, device) to everything the device can give this kernel, so that concurrent callers and
    // processes driving several GPUs through the C ABI can never lower or skip it for one another.
    if (smem > 48 * 1024) {
        static std::atomic<int> smem_granted[kMaxDevices];       // bytes of dynamic shared memory opted in, 0 = not yet
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = -1;
        int granted = dev >= 0 ? smem_granted[dev].load(std::memory_order_acquire) : 0;
        if (granted == 0) {
            int optin = 0;
            cudaFuncAttributes fa;
            cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev < 0 ? 0 : dev);
            if (err == cudaSuccess) err = cudaFuncGetAttributes(&fa, step_kernel<M, NBT, FL>);
            if (err == cudaSuccess) {
                granted = optin - (int)fa.sharedSizeBytes;
                err = cudaFuncSetAttribute(step_kernel<M, NBT, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, granted);
            }
            if (err != cudaSuccess) {
                set_error("step_kernel: cannot opt in to large shared memory: %s", cudaGetErrorString(err));
                return SVBASL_E_CUDA;
            }
            if (dev >= 0) smem_granted[dev].store(granted, std::memory_order_release);
        }
        if ((size_t)granted < smem) {
            set_error("step_kernel needs %zu bytes of shared memory per CTA, the device grants %d", smem, granted);
            return SVBASL_E_UNSUPPORTED;
        }
    }
    step_kernel<M, NBT, FL><<<grid, kBlock, smem, st>>>(a);
    return check_launch("step_kernel");
}

template <class M>
int launch_eval(const EvalArgs &a, cudaStream_t st) {
    const int64_t total = a.n_rows * a.n_batch;
    if (total == 0) return 0;
    eval_kernel<M><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    return check_launch("eval_kernel");
}

template <class M>
int launch_fit(const FitArgs &a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.e.n_vox + kBlock - 1) / kBlock);
    if (grid == 0) return 0;
    fit_kernel<M><<<grid, kBlock, 0, st>>>(a);
    return check_launch("fit_kernel");
}

void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s launch failed: %s", what, cudaGetErrorString(err));
        return SVBASL_E_CUDA;
    }
    return 0;
}

}  // namespace svb
