// model_nn.h - the aslnn surrogate: signal = ftiss * MLP([t, delttiss]), MLP = 2 -> 10 tanh -> 10 tanh -> 1,
// with the forward-mode derivative wrt delttiss carried through the network (what TF autodiff produces for
// /root/reference/svb_models_asl/aslnn.py:93-126, 229-260).  ftiss is LogNormal and delttiss FoldedNormal
// (aslnn.py:73-81), so the engine feeds exp(theta) and |theta|.
//
// SIMT variant: the 151 weights travel in the kernel arguments, i.e. the constant bank, so every FFMA of the
// 10x10 layer reads its weight as an immediate constant operand - no shared-memory or register traffic for them.
// Per element: ~280 FFMA/FMUL + 20 tanh (each MUFU.EX2 + MUFU.RCP): FP32 and XU pipes are about equally loaded.
#pragma once
#include "compat.h"
#include "dev_model.h"

namespace svb {

struct AslNN {
    static constexpr int P = 2;
    static constexpr bool kRegHeavy = true;
    static constexpr int PA = 2;
    static constexpr int H = SVBASL_NN_HIDDEN;

    static constexpr int xf(int p) { return p == 0 ? SVBASL_XF_EXP : SVBASL_XF_ABS; }

    struct Vox {};
    struct Sample {
        float f;
        float a1[H];      // c (W0[1][j] delt + b0[j]), c = 2 log2(e): the delttiss part of layer 1's scaled pre-activation
    };

    template <class Acc>
    static SVB_HD void bind_times(const DevModel &, Vox &, const Acc &) {}
    static SVB_HD Vox load_vox(const DevModel &, int64_t) { return Vox(); }

    static SVB_HD Sample prep_sample(const DevModel &m, const Vox &, const float *x) {
        Sample s;
        s.f = x[0];
        const NNWeights &w = m.nn;
#pragma unroll
        for (int j = 0; j < H; ++j) s.a1[j] = w.w0d_c[j] * x[1] + w.b0_c[j];
        return s;
    }

    // Per row: 10 FFMA + 10 tanh (layer 1), 10 FFMA (1 - h^2), 200 FFMA (the two 10x10 products), 10 tanh, 40 for
    // the output layer and its derivative: ~310 FP32-pipe + 40 MUFU instructions (profiles/r2_notes.md section 4).
    static SVB_HD void eval(const DevModel &m, const Sample &s, float t, float &pred, float *d) {
        const NNWeights &w = m.nn;
        float h1[H], g1[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float h = ftanh_c(w.w0t_c[j] * t + s.a1[j]);
            h1[j] = h;
            g1[j] = 1.0f - h * h;
        }
        float out = w.b2, dout = 0.0f;
#pragma unroll
        for (int k = 0; k < H; ++k) {
            float z = w.b1_c[k], dz = 0.0f;
#pragma unroll
            for (int j = 0; j < H; ++j) {
                z += w.w1_c[k][j] * h1[j];
                dz += w.w1d[k][j] * g1[j];
            }
            const float h = ftanh_c(z);
            out += w.w2[k] * h;
            dout += (w.w2[k] * dz) * (1.0f - h * h);
        }
        pred = s.f * out;
        d[0] = out;
        d[1] = s.f * dout;
    }

    static SVB_HD float predict(const DevModel &m, const Vox &v, const float *x, float t) {
        Sample s = prep_sample(m, v, x);
        float pred, d[PA];
        eval(m, s, t, pred, d);
        return pred;
    }

    template <class Acc>
    static SVB_HD void run(const DevModel &m, const Vox &v, const float *x, Acc &acc) {
        Sample s = prep_sample(m, v, x);
        if (Acc::NB > 0) {
            // not unrolled: one row of the network is ~400 instructions and wants ~60 registers of its own; unrolling
            // the time points only makes the compiler hoist weight loads across rows until it spills
#pragma unroll 1
            for (int b = 0; b < (Acc::NB > 0 ? Acc::NB : 1); ++b) {
                float pred, d[PA];
                eval(m, s, acc.time(b), pred, d);
                acc.add(b, pred, d);
            }
        } else {
            const int nb = acc.n();
            for (int b = 0; b < nb; ++b) {
                float pred, d[PA];
                eval(m, s, acc.time(b), pred, d);
                acc.add(b, pred, d);
            }
        }
    }
};

}  // namespace svb
