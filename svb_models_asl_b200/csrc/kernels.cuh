// kernels.cuh - __global__ wrappers around the per-voxel bodies (voxel_step.h, model_*.h) and their
// launchers.  One thread owns one voxel: every load/store of the SoA arrays is a fully coalesced 128-byte
// row segment per warp, and all per-sample work stays in registers.
#pragma once
#include <cuda_runtime.h>
#include "voxel_step.h"

namespace svb {

constexpr int kBlock = 128;

struct StepArgs {
    DevModel md;
    svbasl_engine e;
    EngineConst ec;
    svbasl_adam ad;
    int32_t update;          // 0: cost + gradient only (svbasl_elbo_grad); 1: fused Adam update (svbasl_step)
    int64_t step;            // iteration index of the first fused iteration (RNG counter / lr_t index)
    float *cost;             // [ld] or NULL
    float *grad;             // [n_state][ld] or NULL
    double *cost_sum;        // [n_iters] or NULL
    long long *nan_count;    // [1] or NULL
};

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
    uint32_t mrfmask;
    int32_t n_params;
    step_launcher_t step;
    eval_launcher_t eval;    // only on the nbt == 0, mrfmask == 0 entry
    fit_launcher_t fit;
};

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

template <class M, int NBT, uint32_t MRFMASK>
__global__ void __launch_bounds__(kBlock) step_kernel(const __grid_constant__ StepArgs a) {
    __shared__ float red[kBlock / 32];
    const int64_t local = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool live = local < a.e.n_vox;
    const int64_t w = a.e.w_begin + (live ? local : 0);
    VoxelStep<M, NBT, MRFMASK> vs;
    vs.load(a.e, w);
    const int n_iters = a.update ? a.ad.n_iters : 1;
    int skipped = 0;
    for (int it = 0; it < n_iters; ++it) {
        const int64_t step = a.step + it;
        const int row0 = (a.update && a.ad.n_batches > 1) ? (int)(step % a.ad.n_batches) : a.e.t_row0;
        float cost = vs.elbo_grad(a.md, a.e, a.ec, w, step, row0);
        if (live) {
            if (a.cost) a.cost[w] = cost;
            if (a.grad) vs.store_grads(a.e, a.grad, w);
            if (a.update) {
                if (vs.grads_finite() && cost == cost) {
                    vs.adam_update(a.e, a.ad, a.ad.lr_t[step], w, it == n_iters - 1);
                } else {
                    ++skipped;
                    if (it == n_iters - 1) vs.store_state(a.e, w);
                    cost = 0.0f;
                }
            }
        } else {
            cost = 0.0f;
        }
        if (a.cost_sum) block_accumulate(cost, a.cost_sum + it, red);
        if (MRFMASK != 0 && a.e.ak_grad) {
#pragma unroll
            for (int k = 0; k < VoxelStep<M, NBT, MRFMASK>::NSP; ++k)
                block_accumulate(live ? vs.ak_out[k] : 0.0f, a.e.ak_grad + k, red);
        }
    }
    if (a.nan_count && skipped) atomicAdd((unsigned long long *)a.nan_count, (unsigned long long)skipped);
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

template <class M, int NBT, uint32_t MRFMASK>
int launch_step(const StepArgs &a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.e.n_vox + kBlock - 1) / kBlock);
    if (grid == 0) return 0;
    step_kernel<M, NBT, MRFMASK><<<grid, kBlock, 0, st>>>(a);
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
