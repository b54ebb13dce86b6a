// compat.h - lets the per-voxel device code (models + fused ELBO/grad/Adam body) also be compiled by
// a plain host compiler.  The host build exists ONLY for tests/hostsim (CPU checks of the kernel
// arithmetic against the oracle in the GPU-less build container); libsvbasl.so never contains it.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SVB_HD __host__ __device__ __forceinline__
#define SVB_D __device__ __forceinline__
#else
#define SVB_HD inline
#define SVB_D inline
#endif

namespace svb {

#if defined(__CUDA_ARCH__)
SVB_D float fexp(float x) { return __expf(x); }            // FMUL + MUFU.EX2
SVB_D float fexp2(float x) { return exp2f(x); }
SVB_D float flog(float x) { return __logf(x); }            // MUFU.LG2 + FMUL
SVB_D float frcp(float x) { return __fdividef(1.0f, x); }   // MUFU.RCP (~1 ulp), no slow path
SVB_D float fdiv(float a, float b) { return __fdividef(a, b); }
SVB_D float fsqrt(float x) {                               // sqrt.approx: MUFU.RSQ/SQRT, no IEEE slow path
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
SVB_D float flog2(float x) { return __log2f(x); }          // MUFU.LG2
SVB_D void fsincos(float x, float *s, float *c) { __sincosf(x, s, c); }     // FMUL.RZ (to turns) + MUFU.SIN + MUFU.COS
SVB_D uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
SVB_D float ferf(float x) { return erff(x); }
SVB_D float ftanh(float x) {                               // 1 - 2/(exp(2x)+1): ~2 ulp, MUFU.EX2 + MUFU.RCP
    float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}
SVB_D bool finite_f(float x) { return isfinite(x); }
// tanh(z) from zc = 2 log2(e) z: 1 - 2 / (2^zc + 1) - MUFU.EX2, FADD, MUFU.RCP, FFMA
SVB_D float ftanh_c(float zc) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zc));
    return fmaf(-2.0f, __fdividef(1.0f, e + 1.0f), 1.0f);
}
#else
SVB_HD float fexp(float x) { return expf(x); }
SVB_HD float fexp2(float x) { return exp2f(x); }
SVB_HD float flog(float x) { return logf(x); }
SVB_HD float frcp(float x) { return 1.0f / x; }
SVB_HD float fdiv(float a, float b) { return a / b; }
SVB_HD float fsqrt(float x) { return sqrtf(x); }
SVB_HD float flog2(float x) { return log2f(x); }
SVB_HD void fsincos(float x, float *s, float *c) {
    *s = sinf(x);
    *c = cosf(x);
}
SVB_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
SVB_HD float ferf(float x) { return erff(x); }
SVB_HD float ftanh(float x) { return tanhf(x); }
SVB_HD bool finite_f(float x) { return isfinite(x); }
SVB_HD float ftanh_c(float zc) { return tanhf(zc * 0.34657359027997264f); }       // zc / (2 log2 e)
#endif

// Wait until this thread's asynchronous global->shared copies (kernels.cuh: Adam moments, neighbour tile) have
// landed.  No-op on the host.
SVB_HD void async_copies_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Warp-level helpers for code with lane-dependent trip counts (model_disp.h).  Every lane of the warp that has not
// exited the kernel must reach them (callers keep the surrounding control flow uniform).  No-ops on the host.
SVB_HD void warp_converge() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}
SVB_HD int warp_max(int v) {
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(0xffffffffu, v);
#else
    return v;
#endif
}

SVB_HD float fmin2(float a, float b) { return a < b ? a : b; }
SVB_HD float fmax2(float a, float b) { return a > b ? a : b; }

}  // namespace svb
