// Micro-benchmark: issue-rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) and MUFU.EX2 on a B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 0.999f, c = 0.001f;
    if (MODE == 0) {            // 8 independent scalar FFMA chains
        for (int i = 0; i < iters; ++i) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    } else if (MODE == 1) {     // the same 8 FMAs as 4 packed FFMA2
        float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
        const float2 mm = make_float2(m, m), cc = make_float2(c, c);
        for (int i = 0; i < iters; ++i) {
            p0 = __ffma2_rn(p0, mm, cc); p1 = __ffma2_rn(p1, mm, cc); p2 = __ffma2_rn(p2, mm, cc); p3 = __ffma2_rn(p3, mm, cc);
        }
        a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
    } else if (MODE == 2) {     // 4 FFMA2 + 4 ALU ops (FMNMX) per iteration: do packed FMAs leave issue slots free?
        float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
        const float2 mm = make_float2(m, m), cc = make_float2(c, c);
        float b0 = seed, b1 = seed + 1, b2 = seed + 2, b3 = seed + 3;
        for (int i = 0; i < iters; ++i) {
            p0 = __ffma2_rn(p0, mm, cc); p1 = __ffma2_rn(p1, mm, cc); p2 = __ffma2_rn(p2, mm, cc); p3 = __ffma2_rn(p3, mm, cc);
            b0 = fminf(b0, p0.x); b1 = fmaxf(b1, p1.x); b2 = fminf(b2, p2.x); b3 = fmaxf(b3, p3.x);
        }
        a0 = p0.x + b0; a1 = p0.y + b1; a2 = p1.x + b2; a3 = p1.y + b3; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
    } else if (MODE == 3) {     // 8 scalar FFMA + 4 ALU ops per iteration
        float b0 = seed, b1 = seed + 1, b2 = seed + 2, b3 = seed + 3;
        for (int i = 0; i < iters; ++i) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
            b0 = fminf(b0, a0); b1 = fmaxf(b1, a2); b2 = fminf(b2, a4); b3 = fmaxf(b3, a6);
        }
        a0 += b0; a1 += b1; a2 += b2; a3 += b3;
    } else {                    // 4: MUFU.EX2 rate
        for (int i = 0; i < iters; ++i) {
            a0 = exp2f(a0 * 0.5f); a1 = exp2f(a1 * 0.5f); a2 = exp2f(a2 * 0.5f); a3 = exp2f(a3 * 0.5f);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int MODE>
float run(float *d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    const double threads = 148.0 * 8 * 256;
    const char *names[] = {"8 FFMA", "4 FFMA2 (=8 FMA)", "4 FFMA2 + 4 FMNMX", "8 FFMA + 4 FMNMX", "4 MUFU.EX2 (+4 FMUL)"};
    float ms[5] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters)};
    const double fmas[5] = {8, 8, 8, 8, 0};
    for (int i = 0; i < 5; ++i)
        printf("%-24s %8.3f ms   %7.2f T FMA/s   %6.2f cycles/iter/SMSP-warp@1.965GHz\n", names[i], ms[i],
               fmas[i] * iters * threads / (ms[i] * 1e-3) / 1e12, ms[i] * 1e-3 * 1.965e9 / iters / (8 * 256 / 32 / 4));
    return 0;
}
