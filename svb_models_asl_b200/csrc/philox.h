// philox.h - counter-based N(0,1) draws for the reparameterised posterior samples.
//
// The reference draws eps with tf.random_normal inside svb (sample = mean +
// chol @ eps, eps [W,P',S]); that stream is not reproducible outside TensorFlow, so parity runs take
// eps from memory and production runs generate it here.  Philox4x32-10 (Salmon et al. 2011), keyed on
// (seed), counter (global voxel id, parameter row j | sample group, step): one call yields the draws of
// parameter row j for 4 consecutive samples, so a voxel's stream does not depend on how voxels are
// sharded over GPUs, and a neighbour's draws can be recomputed instead of exchanged (spatial prior).
#pragma once
#include "compat.h"

namespace svb {

struct Philox4 {
    uint32_t v[4];
};

SVB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

// uniform in (0,1): 24 random bits, never 0 or 1
SVB_HD float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f; }

// 4 standard normals (Box-Muller on two pairs) for (voxel, parameter row j, sample group sg, step)
SVB_HD void normal4(uint64_t seed, int64_t step, int64_t vox_global, int j, int sg, float out[4]) {
    Philox4 r = philox4x32_10((uint32_t)vox_global, (uint32_t)((uint64_t)vox_global >> 32),
                              ((uint32_t)j << 24) | (uint32_t)sg, (uint32_t)step,
                              (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)((uint64_t)step >> 32));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float rad = fsqrt(-2.0f * flog(u01(r.v[2 * h])));
        float s, c;
        fsincos2pi(u01(r.v[2 * h + 1]), &s, &c);
        out[2 * h] = rad * c;
        out[2 * h + 1] = rad * s;
    }
}

}  // namespace svb
