// model_aslrest.h - Buxton general kinetic model for resting-state ASL, forward value and analytic
// derivatives with respect to every inferable parameter, per (voxel, sample, time point).
//
// Takes over AslRestModel.evaluate / tissue_signal / art_signal
// (/root/reference/svb_models_asl/aslrest.py:248-340, 342-391, 393-430) *and* the backward pass that
// TensorFlow autodiff builds from them.  Formulas and derivatives: SURVEY.md Appendix A.1-A.2,
// re-derived with the rate q = 1/T1app as the variable for the infert1 path.
//
// Compile-time layout: the set of inferred parameters is a template bit-mask F (SVBASL_F_* flags), so the
// parameter slots of aslrest.py:183-246 become constant register indices and unused terms vanish.
#pragma once
#include "compat.h"
#include "../../include/svbasl.h"

namespace svb {

// (exp(r u) - 1)/r, finite as r -> 0
SVB_HD float em1r(float r, float u) {
    float z = r * u;
    return fabsf(z) < 1e-4f ? u * (1.0f + 0.5f * z) : expm1f(z) / r;
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
    return (u * (r * e1 + 1.0f) - e1) / r;                            // exp(r u) = r*e1 + 1
}

template <uint32_t F>
struct AslRest {
    static constexpr bool CASL = (F & SVBASL_F_CASL) != 0;
    static constexpr bool ATT = (F & SVBASL_F_INFERATT) != 0;
    static constexpr bool ART = (F & (SVBASL_F_INFERART | SVBASL_F_ARTONLY)) != 0;
    static constexpr bool ARTONLY = (F & SVBASL_F_ARTONLY) != 0;
    static constexpr bool INFWM = (F & SVBASL_F_INFERWM) != 0 && !ARTONLY;
    static constexpr bool INCWM = ((F & SVBASL_F_INCWM) != 0 || INFWM) && !ARTONLY;
    static constexpr bool T1 = (F & SVBASL_F_INFERT1) != 0;
    static constexpr bool TISS = !ARTONLY;

    // parameter order of aslrest.py:183-246
    static constexpr int N_GM = TISS ? (1 + (ATT ? 1 : 0)) : 0;
    static constexpr int I_FTISS = TISS ? 0 : -1;
    static constexpr int I_DELT = (TISS && ATT) ? 1 : -1;
    static constexpr int I_FWM = INFWM ? N_GM : -1;
    static constexpr int I_DELTWM = (INFWM && ATT) ? N_GM + 1 : -1;
    static constexpr int N_TISS = N_GM + (INFWM ? (1 + (ATT ? 1 : 0)) : 0);
    static constexpr int I_T1 = T1 ? N_TISS : -1;
    static constexpr int I_T1WM = (T1 && (F & SVBASL_F_INFERWM)) ? N_TISS + 1 : -1;
    static constexpr int N_A = N_TISS + (T1 ? (1 + ((F & SVBASL_F_INFERWM) ? 1 : 0)) : 0);
    static constexpr int I_FBLOOD = ART ? N_A : -1;
    static constexpr int I_DELTBLOOD = (ART && ATT) ? N_A + 1 : -1;
    static constexpr int P = N_A + (ART ? (1 + (ATT ? 1 : 0)) : 0);

    static constexpr int xf(int) { return SVBASL_XF_IDENTITY; }   // all Normal (aslrest.py:184-246)

    struct Vox {              // per-voxel constants
        float pvgm, pvwm;
    };

    // One tissue compartment's per-sample terms
    struct Tissue {
        float delt, tdp;      // delta, fl(tau + delta)             (mask thresholds, aslrest.py:362-363)
        float q;              // 1/T1app                            (aslrest.py:366)
        float nk;             // -log2(e) * q : E = 2^(nk*(t-delta))
        float Fc;             // CASL: 2*T1app*exp(-delta/t1b)      (aslrest.py:371)
        float c1;             // CASL: exp(tau*q) - 1 ; S_post = Fc*E*c1   (aslrest.py:373, single-exp form)
        float r;              // PASL: r = q - 1/t1b                (aslrest.py:376)
        float erd2;           // PASL: 2*exp(r*delta)
        float e1tau, de1tau;  // PASL: (exp(r tau)-1)/r and its r-derivative (post-bolus, aslrest.py:380)
        float pvf;            // pv * f
        float pv;
        float dqdt1;          // dq/dt1 = -1/t1^2
    };

    struct Sample {
        Tissue gm, wm;
        float fb, deltb, kc, dkc;          // arterial (aslrest.py:404-407)
        float thr_out, ls, inv_ls, dz_in_c, dz_in_t;   // lead-in/out (aslrest.py:411-419)
        bool leadin_ok;
    };

    static SVB_HD Vox load_vox(const svbasl_model &m, int64_t w) {
        Vox v;
        v.pvgm = m.pvgm ? m.pvgm[w] : m.pvgm_s;
        v.pvwm = m.pvwm ? m.pvwm[w] : m.pvwm_s;
        return v;
    }

    static SVB_HD void prep_tissue(const svbasl_model &m, Tissue &ts, float f, float delt, float t1, float pc,
                                   float fcalib, float pv) {
        const float LOG2E = 1.4426950408889634f;
        ts.delt = delt;
        ts.tdp = m.tau + delt;
        float q = frcp(t1) + fcalib / pc;
        ts.q = q;
        ts.nk = -LOG2E * q;
        ts.pv = pv;
        ts.pvf = pv * f;
        ts.dqdt1 = -frcp(t1 * t1);
        float inv_t1b = 1.0f / m.t1b;
        if (CASL) {
            ts.Fc = 2.0f * frcp(q) * fexp(-delt * inv_t1b);
            ts.c1 = fexp(m.tau * q) - 1.0f;
        } else {
            ts.r = q - inv_t1b;
            ts.erd2 = 2.0f * fexp(ts.r * delt);
            ts.e1tau = em1r(ts.r, m.tau);
            ts.de1tau = dem1r(ts.r, m.tau, ts.e1tau);
        }
    }

    // x[P]: model-space parameter values for this sample
    static SVB_HD Sample prep_sample(const svbasl_model &m, const Vox &v, const float *x) {
        Sample s;
        if (TISS) {
            float t1 = T1 ? x[I_T1 < 0 ? 0 : I_T1] : m.t1;
            float delt = ATT ? x[I_DELT < 0 ? 0 : I_DELT] : m.att;
            prep_tissue(m, s.gm, x[I_FTISS < 0 ? 0 : I_FTISS], delt, t1, m.pc, m.fcalib, v.pvgm);
            if (INCWM) {
                float t1wm = (I_T1WM >= 0) ? x[I_T1WM < 0 ? 0 : I_T1WM] : m.t1wm;
                float fwm = INFWM ? x[I_FWM < 0 ? 0 : I_FWM] : m.fwm;
                float dwm = (I_DELTWM >= 0) ? x[I_DELTWM < 0 ? 0 : I_DELTWM] : m.attwm;
                prep_tissue(m, s.wm, fwm, dwm, t1wm, m.pcwm, m.fcalibwm, v.pvwm);
            }
        }
        if (ART) {
            s.fb = x[I_FBLOOD < 0 ? 0 : I_FBLOOD];
            float db = (I_DELTBLOOD >= 0) ? x[I_DELTBLOOD < 0 ? 0 : I_DELTBLOOD] : m.artt;   // Appendix C5
            s.deltb = db;
            float inv_t1b = 1.0f / m.t1b;
            s.kc = CASL ? 2.0f * fexp(-db * inv_t1b) : 0.0f;
            s.dkc = -s.kc * inv_t1b;
            s.thr_out = db + 0.5f * m.tau;
            s.ls = fmin2(db, m.leadscale);
            s.leadin_ok = s.ls > 0.0f;
            float ils = frcp(s.leadin_ok ? s.ls : 1.0f);
            s.inv_ls = ils;
            // d z_in / d deltb = -1/ls when deltb > leadscale; -(t)/deltb^2 when the minimum selects deltb
            bool own = db <= m.leadscale;
            s.dz_in_c = own ? 0.0f : -ils;
            s.dz_in_t = own ? -ils * ils : 0.0f;
        }
        return s;
    }

    // value S (per unit pv*f) and derivatives wrt delta and q of one tissue compartment at time t
    static SVB_HD void tissue_eval(const svbasl_model &m, const Tissue &ts, float t, float &S, float &dSdd,
                                   float &dSdq) {
        bool post = t > ts.tdp;
        bool during = (t > ts.delt) && !post;
        float u = t - ts.delt;
        float inv_t1b = 1.0f / m.t1b;
        if (CASL) {
            float E = fexp2(u * ts.nk);
            float FE = ts.Fc * E;
            float Sd = ts.Fc - FE;
            float Sp = FE * ts.c1;
            float dd = -Sd * inv_t1b - FE * ts.q;
            float dp = Sp * (ts.q - inv_t1b);
            S = post ? Sp : (during ? Sd : 0.0f);
            dSdd = post ? dp : (during ? dd : 0.0f);
            if (T1) {
                float iq = frcp(ts.q);
                float qd = FE * u - Sd * iq;
                float qp = Sp * (m.tau * (ts.c1 + 1.0f) * frcp(ts.c1) - u - iq);
                dSdq = post ? qp : (during ? qd : 0.0f);
            }
        } else {
            // factor*(exp(r t) - exp(r delt)) = 2 exp(-t q) exp(r delt) * (exp(r u) - 1)/r, which (unlike the
            // reference's float32 form) stays accurate when T1app is close to t1b (r -> 0)
            (void)inv_t1b;
            float Be2 = fexp2(t * ts.nk) * ts.erd2;        // 2 exp(-t/T1app) exp(r delt)
            float e1u = em1r(ts.r, u);
            float Sd = Be2 * e1u;
            float Sp = Be2 * ts.e1tau;
            S = post ? Sp : (during ? Sd : 0.0f);
            dSdd = post ? ts.r * Sp : (during ? -Be2 : 0.0f);
            if (T1) {
                float qd = Be2 * dem1r(ts.r, u, e1u) - u * Sd;
                float qp = Be2 * ts.de1tau - u * Sp;
                dSdq = post ? qp : (during ? qd : 0.0f);
            }
        }
    }

    // prediction at time t and d pred / d x[p]
    static SVB_HD void eval(const svbasl_model &m, const Sample &s, float t, float &pred, float *d) {
        pred = 0.0f;
        if (TISS) {
            float S, dd, dq = 0.0f;
            tissue_eval(m, s.gm, t, S, dd, dq);
            pred = s.gm.pvf * S;
            d[I_FTISS < 0 ? 0 : I_FTISS] = s.gm.pv * S;
            if (ATT) d[I_DELT < 0 ? 0 : I_DELT] = s.gm.pvf * dd;
            if (T1) d[I_T1 < 0 ? 0 : I_T1] = s.gm.pvf * dq * s.gm.dqdt1;
            if (INCWM) {
                float Sw, ddw, dqw = 0.0f;
                tissue_eval(m, s.wm, t, Sw, ddw, dqw);
                pred += s.wm.pvf * Sw;
                if (INFWM) d[I_FWM < 0 ? 0 : I_FWM] = s.wm.pv * Sw;
                if (I_DELTWM >= 0) d[I_DELTWM < 0 ? 0 : I_DELTWM] = s.wm.pvf * ddw;
                if (I_T1WM >= 0) d[I_T1WM < 0 ? 0 : I_T1WM] = s.wm.pvf * dqw * s.wm.dqdt1;
            } else if (I_T1WM >= 0) {
                d[I_T1WM < 0 ? 0 : I_T1WM] = 0.0f;
            }
        } else {
            if (T1) d[I_T1 < 0 ? 0 : I_T1] = 0.0f;          // artonly + infert1: parameter exists, unused
            if (I_T1WM >= 0) d[I_T1WM < 0 ? 0 : I_T1WM] = 0.0f;
        }
        if (ART) {
            const float INV_SQRT_PI = 0.5641895835477563f;
            float inv_t1b = 1.0f / m.t1b;
            float inv_LS = 1.0f / m.leadscale;
            float kc = CASL ? s.kc : 2.0f * fexp(-t * inv_t1b);
            float dkc = CASL ? s.dkc : 0.0f;
            bool leadout = t > s.thr_out;
            bool active = leadout || s.leadin_ok;
            float u = t - s.deltb;
            float z = leadout ? -(u - m.tau) * inv_LS : u * s.inv_ls;
            float dz = leadout ? inv_LS : (s.dz_in_c + s.dz_in_t * t);
            float h = 0.5f * (1.0f + ferf(z));
            float g = INV_SQRT_PI * fexp(-z * z);
            float A = active ? kc * h : 0.0f;
            float dA = active ? (dkc * h + kc * g * dz) : 0.0f;
            pred += s.fb * A;
            d[I_FBLOOD < 0 ? 0 : I_FBLOOD] = A;
            if (I_DELTBLOOD >= 0) d[I_DELTBLOOD < 0 ? 0 : I_DELTBLOOD] = s.fb * dA;
        }
    }

    // forward value only (Model.evaluate)
    static SVB_HD float predict(const svbasl_model &m, const Vox &v, const float *x, float t) {
        Sample s = prep_sample(m, v, x);
        float pred, d[P > 0 ? P : 1];
        eval(m, s, t, pred, d);
        return pred;
    }

    // Visit every time point of the batch.  Acc supplies: static NB (compile-time batch size, 0 = dynamic),
    // n(), time(b) and add(b, pred, d).
    template <class Acc>
    static SVB_HD void run(const svbasl_model &m, const Vox &v, const float *x, Acc &acc) {
        Sample s = prep_sample(m, v, x);
        if (Acc::NB > 0) {
#pragma unroll
            for (int b = 0; b < (Acc::NB > 0 ? Acc::NB : 1); ++b) {
                float pred, d[P > 0 ? P : 1];
                eval(m, s, acc.time(b), pred, d);
                acc.add(b, pred, d);
            }
        } else {
            const int nb = acc.n();
            for (int b = 0; b < nb; ++b) {
                float pred, d[P > 0 ? P : 1];
                eval(m, s, acc.time(b), pred, d);
                acc.add(b, pred, d);
            }
        }
    }
};

}  // namespace svb
