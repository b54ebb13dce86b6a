// dev_model.h - the model descriptor as the kernels see it: svbasl_model (include/svbasl.h) with every
// option-only sub-expression folded on the host (reciprocals, T1app rate terms, exp(tau/T1app)-1 ...), so the
// per-element code reads them from the constant bank instead of recomputing or holding them in registers.
#pragma once
#include "compat.h"
#include "../../include/svbasl.h"

namespace svb {

// (exp(r u) - 1)/r, finite as r -> 0
SVB_HD float em1r(float r, float u) {
    float z = r * u;
    return fabsf(z) < 1e-4f ? u * (1.0f + 0.5f * z) : fdiv(expm1f(z), r);
}

// d/dr of em1r(r, u) = (u exp(r u) - em1r)/r; series u^2 (1/2 + z/3 + z^2/8 + ...) where the difference cancels
SVB_HD float dem1r(float r, float u, float e1) {
    float z = r * u;
    if (fabsf(z) < 0.3f) {
        float p = 1.0f / 5040.0f * 7.0f / 8.0f;                       // 7/8! (k = 6)
        p = p * z + 1.0f / 840.0f;
        p = p * z + 1.0f / 144.0f;
        p = p * z + 1.0f / 30.0f;
        p = p * z + 0.125f;
        p = p * z + 1.0f / 3.0f;
        p = p * z + 0.5f;
        return u * u * p;
    }
    return fdiv(u * (r * e1 + 1.0f) - e1, r);                            // exp(r u) = r*e1 + 1
}

// Terms of one tissue compartment that depend only on q = 1/T1app = 1/t1 + fcalib/pc (aslrest.py:366)
struct TissueRates {
    float q, iq;          // 1/T1app, T1app
    float nk;             // -log2(e) * q : exp(-x/T1app) = 2^(nk x)
    float two_iq;         // 2*T1app                                  (CASL factor, aslrest.py:371)
    float c1;             // exp(tau q) - 1 : S_post = F E c1           (aslrest.py:373, single-exp form)
    float tc1;            // tau (c1 + 1)/c1 : d log c1 / dq
    float r;              // q - 1/t1b                                  (PASL, aslrest.py:376)
    float l2e_q_b;        // log2(e) (q - 1/t1b): exp(delta (q - 1/t1b)) = 2^(l2e_q_b delta)
    float e1tau, de1tau;  // (exp(r tau)-1)/r and its r-derivative      (PASL post-bolus, aslrest.py:380)
};

SVB_HD TissueRates tissue_rates(float q, float tau, float inv_t1b, bool casl) {
    TissueRates k;
    k.q = q;
    k.iq = frcp(q);
    k.nk = -1.4426950408889634f * q;
    k.two_iq = 2.0f * k.iq;
    k.l2e_q_b = 1.4426950408889634f * (q - inv_t1b);
    k.c1 = k.tc1 = k.r = k.e1tau = k.de1tau = 0.0f;
    if (casl) {
        k.c1 = fexp(tau * q) - 1.0f;
        k.tc1 = tau * (k.c1 + 1.0f) * frcp(k.c1);
    } else {
        k.r = q - inv_t1b;
        k.e1tau = em1r(k.r, tau);
        k.de1tau = dem1r(k.r, tau, k.e1tau);
    }
    return k;
}

struct NNWeights {                    // aslnn.py:238-240, row-major as in the .npy files
    float w0[2][SVBASL_NN_HIDDEN], b0[SVBASL_NN_HIDDEN];
    float w1[SVBASL_NN_HIDDEN][SVBASL_NN_HIDDEN], b1[SVBASL_NN_HIDDEN];
    float w2[SVBASL_NN_HIDDEN], b2;
    // Folded forms for the fused kernels (model_nn.h).  tanh(z) = 1 - 2 / (2^(c z) + 1) with c = 2 log2(e): the
    // pre-activations are produced already multiplied by c (weights and biases scaled on the host), so a tanh is
    // MUFU.EX2, FADD, MUFU.RCP, FFMA; and the delttiss-derivative of layer 2's pre-activation,
    // sum_j W1[j][k] (1 - h_j^2) W0[1][j], takes its constant factor from w1d[j][k] = W0[1][j] W1[j][k].
    // The two 10x10 tables are stored [k][j] (the FFMA chain of one output unit k reads a contiguous row) in rows of
    // 12 floats on a 16-byte boundary: sm_100 has no constant-bank operand on FFMA, weights reach the FP32 pipe through
    // uniform registers, and only contiguous, aligned constants load four at a time (LDCU.128) - 60 loads per row of
    // the network instead of 200, each of which costs an issue slot (profiles/r2_notes.md section 4).
    float w0t_c[SVBASL_NN_HIDDEN], w0d_c[SVBASL_NN_HIDDEN], b0_c[SVBASL_NN_HIDDEN];      // c W0[0][j], c W0[1][j], c b0[j]
    float b1_c[SVBASL_NN_HIDDEN];                                                          // c b1[k]
    alignas(16) float w1_c[SVBASL_NN_HIDDEN][12];                                          // [k][j] = c W1[j][k]
    alignas(16) float w1d[SVBASL_NN_HIDDEN][12];                                           // [k][j] = W0[1][j] W1[j][k]
};

struct DevModel {
    int32_t kind;
    uint32_t flags;
    float tau, half_tau, inv_t1b;
    float nl2e_inv_t1b;               // -log2(e) / t1b: exp(-x / t1b) = 2^(nl2e_inv_t1b x)
    float att, attwm, fwm, artt;
    float fc_pc, fc_pc_wm;            // fcalib/pc
    float leadscale, inv_leadscale, inv_leadscale_s;   // _s: times sqrt(log2 e)
    float tau_inv_leadscale_s, inv_leadscale_pi;       // tau * inv_leadscale_s; inv_leadscale / sqrt(pi)
    TissueRates gm, wm;               // for the fixed t1 / t1wm of the options
    float pvgm_s, pvwm_s;
    const float *pvgm, *pvwm;
    float conv_dt, conv_tmax;
    int32_t conv_nt;
    float conv_h, conv_inv_h, conv_rho;   // grid step of linspace(0,tmax,nt), its inverse, exp(-h/T1app)
    float s_fixed, sp_fixed;
    NNWeights nn;
};

// host side: fold the options (runs in capi.cu / tests' host build, never per voxel)
inline DevModel make_dev_model(const svbasl_model &m) {
    DevModel d;
    d.kind = m.kind;
    d.flags = m.flags;
    d.tau = m.tau;
    d.half_tau = m.tau / 2;                                            // aslrest.py:411
    d.inv_t1b = 1.0f / m.t1b;
    d.nl2e_inv_t1b = -1.4426950408889634f * d.inv_t1b;
    d.att = m.att;
    d.attwm = m.attwm;
    d.fwm = m.fwm;
    d.artt = m.artt;
    d.fc_pc = m.fcalib / m.pc;
    d.fc_pc_wm = m.pcwm != 0.0f ? m.fcalibwm / m.pcwm : 0.0f;
    d.leadscale = m.leadscale;
    d.inv_leadscale = m.leadscale != 0.0f ? 1.0f / m.leadscale : 0.0f;
    d.inv_leadscale_s = d.inv_leadscale * 1.2011224087864498f;
    d.tau_inv_leadscale_s = m.tau * d.inv_leadscale_s;
    d.inv_leadscale_pi = d.inv_leadscale * 0.5641895835477563f;
    const bool casl = (m.flags & SVBASL_F_CASL) != 0;
    d.gm = tissue_rates((m.t1 > 0.0f ? 1.0f / m.t1 : 0.0f) + d.fc_pc, m.tau, d.inv_t1b, casl);
    d.wm = tissue_rates((m.t1wm > 0.0f ? 1.0f / m.t1wm : 0.0f) + d.fc_pc_wm, m.tau, d.inv_t1b, casl);
    d.pvgm_s = m.pvgm_s;
    d.pvwm_s = m.pvwm_s;
    d.pvgm = m.pvgm;
    d.pvwm = m.pvwm;
    d.conv_dt = m.conv_dt;
    d.conv_tmax = m.conv_tmax;
    d.conv_nt = m.conv_nt;
    d.conv_h = m.conv_nt > 1 ? m.conv_tmax / (float)(m.conv_nt - 1) : 0.0f;
    d.conv_inv_h = d.conv_h > 0.0f ? 1.0f / d.conv_h : 0.0f;
    d.conv_rho = expf(-d.conv_h * d.gm.q);
    d.s_fixed = m.s_fixed;
    d.sp_fixed = m.sp_fixed;
    if (m.kind == SVBASL_MODEL_ASLNN && m.nn_weights) {
        const float *p = m.nn_weights;
        const int H = SVBASL_NN_HIDDEN;
        for (int i = 0; i < 2; ++i) for (int j = 0; j < H; ++j) d.nn.w0[i][j] = *p++;
        for (int j = 0; j < H; ++j) d.nn.b0[j] = *p++;
        for (int i = 0; i < H; ++i) for (int j = 0; j < H; ++j) d.nn.w1[i][j] = *p++;
        for (int j = 0; j < H; ++j) d.nn.b1[j] = *p++;
        for (int j = 0; j < H; ++j) d.nn.w2[j] = *p++;
        d.nn.b2 = *p++;
        const float c = 2.8853900817779268f;                              // 2 log2(e)
        for (int j = 0; j < H; ++j) {
            d.nn.w0t_c[j] = c * d.nn.w0[0][j];
            d.nn.w0d_c[j] = c * d.nn.w0[1][j];
            d.nn.b0_c[j] = c * d.nn.b0[j];
            d.nn.b1_c[j] = c * d.nn.b1[j];
            for (int k = 0; k < H; ++k) {
                d.nn.w1_c[k][j] = c * d.nn.w1[j][k];
                d.nn.w1d[k][j] = d.nn.w0[1][j] * d.nn.w1[j][k];
            }
        }
    } else {
        d.nn = NNWeights();
    }
    return d;
}

}  // namespace svb
