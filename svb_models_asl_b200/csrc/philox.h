// philox.h - counter-based N(0,1) draws for the reparameterised posterior samples.
//
// The reference draws eps with tf.random_normal inside svb (sample = mean +
// chol @ eps, eps [W,P',S]); that stream is not reproducible outside TensorFlow, so parity runs take
// eps from memory and production runs generate it here.  Philox2x32-10 (Salmon et al. 2011, Random123):
// counter = (global voxel id, pair number), key = mix(seed, step).  One call yields a Box-Muller pair = two
// consecutive normals of the voxel's stream for that step (numbering below); a voxel's stream depends only on
// (seed, step, global voxel id): results do not depend on how voxels are sharded over GPUs.
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

// Box-Muller pair from two 32-bit words in 12 instructions: u = (x + 1/2) 2^-32 straight from the unsigned-to-float
// conversion (one I2FP + one FFMA per word; the conversion rounds to nearest, so u lies in [2^-33, 1] and the
// radius in [0, 6.8]), the radius as sqrt(-2 ln2 * lg2 u) (MUFU.LG2, FMUL, MUFU.SQRT) and the angle with 2 pi
// folded into the conversion's FFMA.
SVB_HD void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float ua = (float)a * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
    const float ang = (float)b * (6.283185307179586f * 2.3283064365386963e-10f) + (6.283185307179586f * 1.1641532182693481e-10f);
    const float rad = fsqrt(-1.3862943611198906f * flog2(ua));
    float sn, cs;
    fsincos(ang, &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

// The stream of one (voxel, step): normals numbered q = j*S2 + s (posterior row j, sample s; S2 = S rounded up to
// even).  Normals 2p and 2p+1 are the Box-Muller pair of ONE Philox call with counter (global voxel, p): a call
// serves the SAME row of two consecutive samples.  The sample loop therefore draws all N rows for samples s and s+1
// together on even s (N calls, N/2 per sample whatever N is), and one row of all S samples - what a neighbour needs
// of a spatially regularised parameter - costs S/2 calls.
SVB_HD void normal_pair(uint32_t key, int64_t vox_global, int pair, float &n0, float &n1) {
    uint32_t a, b;
    philox2x32_10((uint32_t)vox_global, (uint32_t)pair ^ ((uint32_t)((uint64_t)vox_global >> 32) << 24), key, a, b);
    box_muller(a, b, n0, n1);
}

SVB_HD int stream_pair(int j, int s, int S) { return (j * ((S + 1) >> 1)) + (s >> 1); }

// one normal: row j of sample s (slow path: single rows, tests)
SVB_HD float normal_at(uint32_t key, int64_t vox_global, int j, int s, int S) {
    float n0, n1;
    normal_pair(key, vox_global, stream_pair(j, s, S), n0, n1);
    return (s & 1) ? n1 : n0;
}

// All N draws of sample s inside a loop over s = 0, 1, 2, ...: an even sample draws the pairs and leaves the second
// halves in `spare` for the odd sample that follows.  The branch on s is uniform across the warp.
template <int N>
SVB_HD void normal_row(uint32_t key, int64_t vox_global, int s, int S, float *eps, float *spare) {
    if ((s & 1) == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) normal_pair(key, vox_global, stream_pair(j, s, S), eps[j], spare[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) eps[j] = spare[j];
    }
}

}  // namespace svb
