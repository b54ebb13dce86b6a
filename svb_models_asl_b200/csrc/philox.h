// philox.h - counter-based N(0,1) draws for the reparameterised posterior samples.
//
// The reference draws eps with tf.random_normal inside svb (sample = mean +
// chol @ eps, eps [W,P',S]); that stream is not reproducible outside TensorFlow, so parity runs take
// eps from memory and production runs generate it here.  Philox2x32-10 (Salmon et al. 2011, Random123):
// counter = (global voxel id, sample index | parameter pair), key = mix(seed, step).  One call yields the
// draws of parameter rows 2k and 2k+1 for one sample (a Box-Muller pair), so nothing has to be cached across
// samples, and a voxel's stream depends only on (seed, step, global voxel id): results do not depend on how
// voxels are sharded over GPUs.
#pragma once
#include "compat.h"

namespace svb {

SVB_HD void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t &o0, uint32_t &o1) {
    const uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t prod = (uint64_t)M * (uint64_t)c0;     // one IMAD.WIDE.U32
        c0 = (uint32_t)(prod >> 32) ^ key ^ c1;
        c1 = (uint32_t)prod;
        key += W;
    }
    o0 = c0;
    o1 = c1;
}

SVB_HD uint32_t rng_key(uint64_t seed, int64_t step) {
    // odd multipliers: for a fixed seed distinct steps (mod 2^32) give distinct keys
    return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA77u) ^ ((uint32_t)step * 0x9E3779B1u) ^
           ((uint32_t)((uint64_t)step >> 32) * 0xC2B2AE3Du);
}

// uniform in (0,1): 24 random bits, never 0 or 1
SVB_HD float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f; }

// two standard normals for (voxel, sample s, parameter pair k): rows 2k and 2k+1
SVB_HD void normal2(uint32_t key, int64_t vox_global, int s, int pair, float &n0, float &n1) {
    uint32_t a, b;
    philox2x32_10((uint32_t)vox_global, ((uint32_t)s << 4 | (uint32_t)pair) ^ ((uint32_t)((uint64_t)vox_global >> 32) << 24),
                  key, a, b);
    const float rad = fsqrt(-2.0f * flog(u01(a)));
    float sn, cs;
    fsincos2pi(u01(b), &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

// all N draws of one sample
template <int N>
SVB_HD void normal_row(uint32_t key, int64_t vox_global, int s, float *eps) {
#pragma unroll
    for (int k = 0; k < (N + 1) / 2; ++k) {
        float n0, n1;
        normal2(key, vox_global, s, k, n0, n1);
        eps[2 * k] = n0;
        if (2 * k + 1 < N) eps[2 * k + 1] = n1;
    }
}

}  // namespace svb
