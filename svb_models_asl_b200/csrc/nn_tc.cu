// nn_tc.cu - the aslnn surrogate's batched hidden-layer GEMM on the 5th-generation tensor cores.
//
// AslNNModel.evaluate (/root/reference/svb_models_asl/aslnn.py:93-126, 229-260) is, per row (voxel, sample, time
// point), signal = ftiss * (W2 . tanh(W1^T tanh(W0^T [t, delt] + b0) + b1) + b2).  The 2->10 and 10->1 layers are
// a handful of FMAs and stay on the FP32 pipe; the 10x10 layer is a [rows x 10] . [10 x 10] GEMM, done here as
//   D[128 x 16] (TMEM, fp32) = A[128 x 32] (smem, tf32) . B[32 x 16] (smem, tf32)       tcgen05.mma kind::tf32
// per tile of 128 rows.  Accuracy: tf32 keeps 10 mantissa bits, far short of the 1e-5 forward tolerance, so every
// operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, lo = x - hi) and the product is
// formed as hi.Whi + lo.Whi + hi.Wlo by concatenating along K:  A = [hi(10) | lo(10) | hi(10) | 0 0],
// B = [Whi; Whi; Wlo; 0 0]  (K = 32 = four K=8 MMAs), error ~2^-20 relative.
//
// Data movement: the pre-arranged B tile (2 KB, canonical K-major no-swizzle core-matrix layout) arrives by one
// TMA bulk copy (cp.async.bulk -> mbarrier complete_tx); each thread writes its row of A with 8 STS.128 (conflict
// free: a core matrix is 8 rows x 16 B, consecutive rows are consecutive 16-B words); one elected thread issues
// the MMAs and tcgen05.commit; every thread then pulls its own accumulator row with tcgen05.ld 32x32b.x16
// (TMEM lane == tile row) and finishes tanh + the output layer in registers.
//
// All waits are bounded: a descriptor mistake must surface as an error code, never as a hung GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "dev_model.h"
#include "compat.h"

namespace svb {

void set_error(const char *fmt, ...);

constexpr int kTcRows = 128;          // M
constexpr int kTcN = 16;              // N (10 used)
constexpr int kTcK = 32;              // concatenated K (30 used)
constexpr int kTcChunks = kTcK / 4;   // 16-byte K chunks (4 tf32 each)
constexpr uint32_t kALbo = kTcRows * 16;      // bytes between K-adjacent core matrices of A (one chunk plane)
constexpr uint32_t kASbo = 128;               // bytes between M-adjacent core matrices (8 rows x 16 B)
constexpr uint32_t kBLbo = kTcN * 16;
constexpr uint32_t kBSbo = 128;
constexpr int kBTileBytes = kTcChunks * kTcN * 16;   // 2048
constexpr int kATileBytes = kTcChunks * kTcRows * 16;   // 16384
constexpr uint32_t kTmemCols = 32;

struct NnTcArgs {
    NNWeights w;                      // layer 0 / biases / layer 2 from the constant bank
    const float *b_tile;              // device: B operand tile, kBTileBytes, layout of b_tile_index()
    const float *params;              // [2][n_rows]: ftiss, delttiss
    const float *tpts;                // [n_t_rows][B]
    float *out;                       // [n_rows][B] or nullptr (then `hidden` is written)
    float *hidden;                    // optional [n_elems][10]: second-layer pre-activations (tests)
    int64_t n_rows, n_t_rows;         // rows of params; out has n_rows*B elements
    int32_t n_batch;
    int32_t *status;                  // device: set non-zero when a bounded wait expires
};

// element (n, k) of the K-major no-swizzle B tile, in floats
__host__ __device__ inline int b_tile_index(int n, int k) {
    return ((k >> 2) * (int)kBLbo + (n >> 3) * (int)kBSbo + (n & 7) * 16 + (k & 3) * 4) / 4;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SM100 shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout type = 0
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) |
                            ((uint32_t)(kTcRows >> 4) << 24);

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(kTcRows) nn_eval_tc_kernel(const __grid_constant__ NnTcArgs a) {
    __shared__ __align__(1024) float a_tile[kATileBytes / 4];
    __shared__ __align__(128) float b_tile[kBTileBytes / 4];
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_tma_u = smem_u32(&bar_tma), bar_mma_u = smem_u32(&bar_mma);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_tma_u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mma_u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // TMA bulk copy of the weight tile; completion is signalled on bar_tma by byte count
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_tma_u), "r"((uint32_t)kBTileBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(b_tile)), "l"(a.b_tile), "r"((uint32_t)kBTileBytes), "r"(bar_tma_u) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    bool ok = mbar_wait_bounded(bar_tma_u, 0);

    const uint64_t adesc0 = umma_desc(smem_u32(a_tile), kALbo, kASbo);
    const uint64_t bdesc0 = umma_desc(smem_u32(b_tile), kBLbo, kBSbo);
    const int64_t n_elems = a.n_rows * a.n_batch;
    const int64_t n_tiles = (n_elems + kTcRows - 1) / kTcRows;
    const int64_t rows_per_t = a.n_rows / a.n_t_rows;
    uint32_t phase = 0;
    const int H = SVBASL_NN_HIDDEN;

    for (int64_t tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int64_t idx = tile * kTcRows + tid;
        const bool live = idx < n_elems;
        const int64_t row = live ? idx / a.n_batch : 0;
        const int b = live ? (int)(idx - row * a.n_batch) : 0;
        const float t = a.tpts[(row / rows_per_t) * a.n_batch + b];
        const float f = a.params[row];
        const float delt = a.params[a.n_rows + row];
        // layer 0 on the FP32 pipe, then the hi / lo / hi split rows of A
        float hi[H], lo[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float h = ftanh(a.w.w0[0][j] * t + a.w.w0[1][j] * delt + a.w.b0[j]);
            const float hh = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);
            hi[j] = hh;
            lo[j] = h - hh;
        }
        float arow[kTcK];
#pragma unroll
        for (int j = 0; j < H; ++j) { arow[j] = hi[j]; arow[H + j] = lo[j]; arow[2 * H + j] = hi[j]; }
        arow[30] = arow[31] = 0.0f;
        float4 *dst = reinterpret_cast<float4 *>(a_tile) + (tid >> 3) * (kASbo / 16) + (tid & 7);
#pragma unroll
        for (int c = 0; c < kTcChunks; ++c)
            dst[c * (kALbo / 16)] = make_float4(arow[4 * c], arow[4 * c + 1], arow[4 * c + 2], arow[4 * c + 3]);
        // generic-proxy writes -> visible to the tensor core's async proxy, then hand over to the issuing thread
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kTcK / 8; ++ks) {
                // each K=8 slice spans two 16-byte chunks: advance the start address by 2*LBO (in 16-B units)
                const uint64_t ad = adesc0 + (uint64_t)((2 * ks * kALbo) >> 4);
                const uint64_t bd = bdesc0 + (uint64_t)((2 * ks * kBLbo) >> 4);
                const uint32_t accum = ks > 0 ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(kIdesc), "r"(accum) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma_u) : "memory");
        }
        ok = mbar_wait_bounded(bar_mma_u, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[16];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float acc = a.w.b2;
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const float z = __uint_as_float(r[k]) + a.w.b1[k];
            if (a.hidden && live) a.hidden[idx * H + k] = z;
            acc += a.w.w2[k] * ftanh(z);
        }
        if (a.out && live) a.out[idx] = f * acc;
        // D and the A tile are reused by the next tile: everyone must be done reading first
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (!ok && a.status) atomicExch(a.status, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols));
    }
}

// Host: B operand tile for weights W1 [in j][out k] (aslnn.py:239): row n = output unit, K index = concatenated input
void nn_tc_build_b_tile(const NNWeights &w, float *tile /* kBTileBytes/4 floats */) {
    for (int i = 0; i < kBTileBytes / 4; ++i) tile[i] = 0.0f;
    const int H = SVBASL_NN_HIDDEN;
    for (int n = 0; n < H; ++n) {
        for (int j = 0; j < H; ++j) {
            const float v = w.w1[j][n];
            uint32_t bits;
            memcpy(&bits, &v, 4);
            bits &= 0xFFFFE000u;
            float hi;
            memcpy(&hi, &bits, 4);
            const float lo = v - hi;
            tile[b_tile_index(n, j)] = hi;             // multiplies A's hi part
            tile[b_tile_index(n, H + j)] = hi;         // multiplies A's lo part
            tile[b_tile_index(n, 2 * H + j)] = lo;     // multiplies A's (second) hi part
        }
    }
}

static int launch_nn_eval_tc(const NnTcArgs &a, cudaStream_t st);

int nn_tc_launch(const DevModel &dm, const float *b_tile, const float *params, const float *tpts, float *out,
                 float *hidden, int64_t n_rows, int32_t n_batch, int64_t n_t_rows, int32_t *status, cudaStream_t st) {
    NnTcArgs a;
    a.w = dm.nn;
    a.b_tile = b_tile;
    a.params = params;
    a.tpts = tpts;
    a.out = out;
    a.hidden = hidden;
    a.n_rows = n_rows;
    a.n_t_rows = n_t_rows;
    a.n_batch = n_batch;
    a.status = status;
    return launch_nn_eval_tc(a, st);
}

static int launch_nn_eval_tc(const NnTcArgs &a, cudaStream_t st) {
    const int64_t n_tiles = (a.n_rows * a.n_batch + kTcRows - 1) / kTcRows;
    if (n_tiles == 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t grid = n_tiles < (int64_t)sms * 8 ? n_tiles : (int64_t)sms * 8;
    nn_eval_tc_kernel<<<(unsigned)grid, kTcRows, 0, st>>>(a);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("nn_eval_tc_kernel launch failed: %s", cudaGetErrorString(err));
        return SVBASL_E_CUDA;
    }
    return 0;
}

}  // namespace svb
