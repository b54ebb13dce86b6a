// model_nn_tc.cuh - the aslnn surrogate INSIDE the fused SVB step with its two 10x10 products on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM).
//
// AslNNModel.evaluate (/root/reference/svb_models_asl/aslnn.py:93-126, 229-260) per row (voxel, sample, time point):
//   h1 = tanh(W0^T [t, delt] + b0),  z2 = W1^T h1 + b1,  out = W2 . tanh(z2) + b2,  signal = ftiss * out
// and, for the gradient with respect to delttiss (forward mode, what TF autodiff yields):
//   dz2 = W1^T ((1 - h1^2) * W0[1]),  dout = W2 . ((1 - tanh(z2)^2) * dz2).
// The FP32-pipe kernel (model_nn.h) spends 200 of its ~410 instructions per row on the two 10x10 products plus ~50
// uniform loads of their weights (profiles/r2_notes.md section 4).  Here one thread still owns one voxel, so the 128
// threads of a CTA hold 128 rows at a time = one M = 128 tile:
//   D1[128 x 16] = A1[128 x 32] . B1[32 x 16]     A1 = [hi(h1) | hi(h1) | lo(h1) | 1 1],  B1 = [Whi; Wlo; Whi; bias_hi; bias_lo]
//   D2[128 x 16] = A2[128 x 32] . B2[32 x 16]     A2 = the same split of (1 - h1^2),      B2 from W0[1][j] W1[j][k] W2[k]
// kind::f16 with float32 accumulation; every operand is split hi + lo into two fp16 values (22 significant bits) and
// the three leading products are kept, so the 10-term sums carry ~1e-7 relative error - float32 level.  W and the bias
// arrive already multiplied by 2 log2(e), so D1 feeds ftanh_c directly.
//
// The A operands live in TENSOR MEMORY, not in shared memory: row i of an M = 128 operand is TMEM lane i, so a thread
// writes its own A rows with tcgen05.st (32 halves = 16 columns each) straight from registers.  (The first version
// staged A in shared memory as tf32: 16 STS.128 per thread and row plus the tensor core's reads of the same 32 KB made
// the shared-memory pipe the limiter - profiles/r2_notes.md section 4.)  Only the two 1 KB weight tiles sit in shared
// memory.  The cost that remains is the rendezvous - a row's MMAs need all 128 rows - so a round carries TWO time
// points: four tcgen05.st per thread, one CTA barrier, one thread (the duty rotates over the warps) issues eight
// K = 16 MMAs and a commit, every thread waits on the mbarrier and pulls 32 accumulator columns per row with
// tcgen05.ld.  TMEM columns per CTA: 2 x (A1 16 | A2 16) + 2 x (D1 16 | D2 16) = 128, so four CTAs share an SM's 512
// columns and one CTA's tensor-core round trip is covered by the others' FP32 / MUFU work.
//
// All waits are bounded; an expired wait raises svbasl_engine-independent status (NnTcShared::failed) that makes the
// kernel skip its updates - a descriptor mistake must never hang the GPU.
#pragma once
#include "model_nn.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace svb {

namespace nntc {
constexpr int kRows = 128, kN = 16, kK = 32, kChunks = kK / 8;      // K = 32 halves = four 16-byte chunks
constexpr int kRowsPerRound = 2;                       // time points per rendezvous
constexpr uint32_t kBLbo = kN * 16, kBSbo = 128;
constexpr int kBTileHalves = kChunks * kN * 8;        // 512 halves = 1 KB
constexpr uint32_t kColsA = kK / 2;                   // an A row of 32 halves = 16 TMEM columns
// TMEM columns of a CTA: per row slot r: [A1 16 | A2 16] at 32 r, accumulators [D1 16 | D2 16] at 64 + 32 r
constexpr uint32_t kTmemCols = 128;
constexpr uint32_t kColD = 2 * kColsA * kRowsPerRound;
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);

__host__ __device__ inline int b_index(int n, int k) {   // element (n, k) of the K-major no-swizzle B tile, in halves
    return ((k >> 3) * (int)kBLbo + (n >> 3) * (int)kBSbo + (n & 7) * 16 + (k & 7) * 2) / 2;
}
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// the leading 11 significant bits of x: exactly representable in fp16 (for |x| >= 2^-14; below that the conversion
// rounds and the error is below 2^-25 absolute)
__device__ __forceinline__ float f16_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// two tanh from ONE reciprocal: 1/a0 = a1/(a0 a1), 1/a1 = a0/(a0 a1).  The tensor-core step is bound by the XU pipe
// (40 MUFU per row for 20 tanh, ncu: 73 % busy, mio_throttle the top stall) while the FP32 pipe idles at 30 %: this
// trades a quarter of the MUFU work for 4 FP32/ALU instructions per pair.  2^zc is capped at 2^62 so that the product
// of two denominators stays finite (tanh is 1 to float32 precision long before).
__device__ __forceinline__ void tanh_pair_c(float zc0, float zc1, float &t0, float &t1) {
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(zc0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(zc1));
    const float a0 = fminf(e0, 4.611686018427388e18f) + 1.0f, a1 = fminf(e1, 4.611686018427388e18f) + 1.0f;
    const float r = __fdividef(1.0f, a0 * a1);
    t0 = fmaf(-2.0f, r * a1, 1.0f);
    t1 = fmaf(-2.0f, r * a0, 1.0f);
}
__device__ __forceinline__ uint32_t pack2(float e0, float e1) {      // e0 in the low half = the lower K index
    const __half2 h = __floats2half2_rn(e0, e1);
    return *reinterpret_cast<const uint32_t *>(&h);
}
#endif
}  // namespace nntc

#if defined(__CUDACC__)
struct AslNNTC : AslNN {
    static constexpr bool kCtaCoop = true;
    static constexpr bool kRegHeavy = false;       // 4 CTAs per SM: the others cover one CTA's tensor-core round trip

    struct Vox {
        uint32_t bar, tmem, phase;    // mbarrier (shared-window address), TMEM base, (rounds issued << 1) | parity of the next completion
        uint32_t tlane;               // TMEM address of this warp's 32 lanes, column 0 of the allocation
        uint64_t bd1, bd2;            // shared-memory matrix descriptors of the weight tiles
        int *failed;
    };

    struct Shared {
        alignas(128) __half b1[nntc::kBTileHalves];
        alignas(128) __half b2[nntc::kBTileHalves];
        alignas(8) uint64_t bar;
        uint32_t tmem_base;
        int failed;
        uint32_t phase[nntc::kRows];  // per thread: the phase word (survives load_vox)
    };

    static __device__ __forceinline__ Shared &shared() {
        __shared__ Shared sh;
        return sh;
    }

    // once per CTA: weight tiles (hi / lo split, bias rows), mbarrier, TMEM columns
    static __device__ void cta_begin(const DevModel &m) {
        Shared &sh = shared();
        const NNWeights &w = m.nn;
        const int tid = threadIdx.x;
        for (int i = tid; i < nntc::kBTileHalves; i += blockDim.x) { sh.b1[i] = __float2half(0.0f); sh.b2[i] = __float2half(0.0f); }
        __syncthreads();
        for (int i = tid; i < H * H; i += blockDim.x) {
            const int n = i / H, j = i - n * H;                    // n = output unit k, j = input unit
            const float v1 = w.w1_c[n][j], v2 = w.w1d[n][j] * w.w2[n];       // D2 arrives multiplied by the output weight
            const __half h1 = __float2half_rn(v1), h2 = __float2half_rn(v2);
            const __half l1 = __float2half_rn(v1 - __half2float(h1)), l2 = __float2half_rn(v2 - __half2float(h2));
            sh.b1[nntc::b_index(n, j)] = h1;                       // against A's hi
            sh.b1[nntc::b_index(n, H + j)] = l1;                   // against A's hi (second copy)
            sh.b1[nntc::b_index(n, 2 * H + j)] = h1;               // against A's lo
            sh.b2[nntc::b_index(n, j)] = h2;
            sh.b2[nntc::b_index(n, H + j)] = l2;
            sh.b2[nntc::b_index(n, 2 * H + j)] = h2;
        }
        if (tid < H) {                                             // bias rows: A1 carries 1.0 in K = 30, 31
            const float b = w.b1_c[tid];
            const __half bh = __float2half_rn(b);
            sh.b1[nntc::b_index(tid, 30)] = bh;
            sh.b1[nntc::b_index(tid, 31)] = __float2half_rn(b - __half2float(bh));
        }
        sh.phase[tid & (nntc::kRows - 1)] = 0u;
        if (tid == 0) {
            sh.failed = 0;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(nntc::smem_u32(&sh.bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if ((tid >> 5) == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(nntc::smem_u32(&sh.tmem_base)), "r"(nntc::kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // B tiles: generic-proxy writes -> async proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    static __device__ void cta_end() {
        Shared &sh = shared();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if ((threadIdx.x >> 5) == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "r"(nntc::kTmemCols));
    }
    static __device__ bool cta_failed() { return shared().failed != 0; }

    template <class Acc>
    static __device__ __forceinline__ void bind_times(const DevModel &, Vox &, const Acc &) {}

    static __device__ __forceinline__ Vox load_vox(const DevModel &, int64_t) {
        Shared &sh = shared();
        Vox v;
        const int tid = threadIdx.x;
        v.bar = nntc::smem_u32(&sh.bar);
        v.tmem = sh.tmem_base;
        v.tlane = sh.tmem_base + ((uint32_t)((tid >> 5) * 32) << 16);
        v.phase = sh.phase[tid];
        v.bd1 = nntc::umma_desc(nntc::smem_u32(sh.b1), nntc::kBLbo, nntc::kBSbo);
        v.bd2 = nntc::umma_desc(nntc::smem_u32(sh.b2), nntc::kBLbo, nntc::kBSbo);
        v.failed = &sh.failed;
        return v;
    }

    // one A row of 32 halves [hi(10) | hi(10) | lo(10) | tail tail] -> this thread's TMEM lane, 16 columns from `taddr`
    // (hi + lo = x to 2^-22: with the weights split the same way the three kept products hi*Whi + hi*Wlo + lo*Whi give
    // the 10-term sums to ~1e-7, accumulated in float32)
    static __device__ __forceinline__ void store_row(uint32_t taddr, const float *x, uint32_t tail2) {
        uint32_t r[nntc::kColsA];
#pragma unroll
        for (int j = 0; j < H / 2; ++j) {
            const float h0 = nntc::f16_hi(x[2 * j]), h1 = nntc::f16_hi(x[2 * j + 1]);
            r[j] = nntc::pack2(h0, h1);
            r[H / 2 + j] = r[j];
            r[H + j] = nntc::pack2(x[2 * j] - h0, x[2 * j + 1] - h1);
        }
        r[15] = tail2;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
            ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
              "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
            : "memory");
    }

    static __device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        return ok != 0;
    }
    // bounded: gives up when another thread of the CTA has, or after ~4 M polls of its own
    static __device__ __noinline__ bool wait_slow(uint32_t bar, uint32_t parity, int *failed) {
        for (int spin = 0; spin < (1 << 22); ++spin) {
            if (try_wait(bar, parity)) return true;
            if (*(volatile int *)failed != 0) return false;
        }
        atomicExch(failed, 1);
        return false;
    }

    // layer 1 of one row and its two A rows -> row slot `slot` of tensor memory
    static __device__ __forceinline__ void stage_row(const NNWeights &w, const Sample &s, const Vox &v, float t, int slot) {
        float h1[H], g1[H];
#pragma unroll
        for (int j = 0; j < H; j += 2) {
            nntc::tanh_pair_c(w.w0t_c[j] * t + s.a1[j], w.w0t_c[j + 1] * t + s.a1[j + 1], h1[j], h1[j + 1]);
            g1[j] = 1.0f - h1[j] * h1[j];
            g1[j + 1] = 1.0f - h1[j + 1] * h1[j + 1];
        }
        const uint32_t base = v.tlane + (uint32_t)slot * 2u * nntc::kColsA;
        store_row(base, h1, 0x3C003C00u);                          // (1, 1): the bias columns
        store_row(base + nntc::kColsA, g1, 0u);
    }

    // the staged rows are handed to the tensor core: CTA barrier, 4 MMAs per row (2 tiles x K = 32 in two K = 16 steps), commit
    static __device__ __forceinline__ void issue(Vox &v, int nrows) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");      // the rows are in TMEM: hand over to the issuer
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // the issuing duty goes round the four warps (= the SM's four schedulers) round by round: a fixed issuer would
        // load one scheduler with every resident CTA's issue work
        const uint32_t issuer = ((v.phase >> 1) & 3u) << 5;
        v.phase += 2u;
        if (threadIdx.x == issuer) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int slot = 0; slot < nntc::kRowsPerRound; ++slot) {
                if (slot < nrows) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t bd0 = half ? v.bd2 : v.bd1;
                        const uint32_t a0 = v.tmem + (uint32_t)slot * 2u * nntc::kColsA + (half ? nntc::kColsA : 0u);
                        const uint32_t d = v.tmem + nntc::kColD + (uint32_t)(slot * 2 * nntc::kN + half * nntc::kN);
#pragma unroll
                        for (int ks = 0; ks < nntc::kK / 16; ++ks) {
                            // a K = 16 slice of A is 8 TMEM columns; of B two 16-byte chunks: start address + 2*LBO
                            const uint32_t at = a0 + (uint32_t)(8 * ks);
                            const uint64_t bd = bd0 + (uint64_t)((2 * ks * nntc::kBLbo) >> 4);
                            const uint32_t accum = ks > 0 ? 1u : 0u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                ::"r"(d), "r"(at), "l"(bd), "r"(nntc::kIdesc), "r"(accum) : "memory");
                        }
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(v.bar) : "memory");
        }
    }

    // accumulators of row slot `slot` -> layer 2, output layer, residual sums
    template <class Acc>
    static __device__ __forceinline__ void finish_row(const NNWeights &w, const Sample &s, const Vox &v, int slot, int b, Acc &acc) {
        uint32_t r[32];
        const uint32_t taddr = v.tlane + nntc::kColD + (uint32_t)(slot * 2 * nntc::kN);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float out = w.b2, dout = 0.0f;
#pragma unroll
        for (int k = 0; k < H; k += 2) {
            float ha, hb;
            nntc::tanh_pair_c(__uint_as_float(r[k]), __uint_as_float(r[k + 1]), ha, hb);   // D1 = 2 log2(e) (W1^T h1 + b1)
            out += w.w2[k] * ha;
            out += w.w2[k + 1] * hb;
            dout += __uint_as_float(r[nntc::kN + k]) * (1.0f - ha * ha);    // D2 = W2[k] * dz2[k]
            dout += __uint_as_float(r[nntc::kN + k + 1]) * (1.0f - hb * hb);
        }
        float d[PA];
        d[0] = out;
        d[1] = s.f * dout;
        acc.add(b, s.f * out, d);
    }

    // Every thread of the CTA calls this together (the sample and time-point loops are uniform).  Two time points per
    // round: stage both rows, one rendezvous, eight MMAs, one wait, both rows' layer 2.  Nothing of a CTA overlaps its
    // own tensor-core round trip - the three other CTAs of the SM do.
    template <class Acc>
    static __device__ __forceinline__ void run(const DevModel &m, Vox &v, const float *x, Acc &acc) {
        const NNWeights &w = m.nn;
        const Sample s = AslNN::prep_sample(m, AslNN::Vox(), x);
        constexpr bool kStatic = Acc::NB > 0;
        const int nb = kStatic ? Acc::NB : acc.n();
#pragma unroll
        for (int b = 0; b < (kStatic ? Acc::NB : 1 << 30); b += nntc::kRowsPerRound) {
            if (!kStatic && b >= nb) break;
            const int nrows = (nb - b) < nntc::kRowsPerRound ? (nb - b) : nntc::kRowsPerRound;
#pragma unroll
            for (int slot = 0; slot < nntc::kRowsPerRound; ++slot)
                if (slot < nrows) stage_row(w, s, v, acc.time(b + slot), slot);
            issue(v, nrows);
            if (!try_wait(v.bar, v.phase & 1u)) wait_slow(v.bar, v.phase & 1u, v.failed);
            v.phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int slot = 0; slot < nntc::kRowsPerRound; ++slot)
                if (slot < nrows) finish_row(w, s, v, slot, b + slot, acc);
            // (the next round's tcgen05.st overwrite the A slots - their MMAs have completed - and its MMAs, issued
            // behind the next CTA barrier, the accumulators every thread has finished loading by then)
        }
        shared().phase[threadIdx.x] = v.phase;
    }
};
#endif

}  // namespace svb
