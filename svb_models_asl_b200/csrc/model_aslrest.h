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
// Everything that depends only on the options (rates, exp(tau/T1app)-1, reciprocals) is folded on the host
// into DevModel (dev_model.h) and read straight from the constant bank.
#pragma once
#include "compat.h"
#include "dev_model.h"

namespace svb {

// 1/2 (1 + erf z) and exp(-z^2) together, branch-free, from ONE exponential:
// erfc(|z|) = (a1 t + ... + a5 t^5) exp(-z^2), t = 1/(1 + p|z|)  (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7),
// so the smoothed step is accurate to 7.5e-8 absolute - below float32 resolution of values near 1.
// The argument arrives pre-scaled, zs = z*sqrt(log2 e), so that exp(-z^2) = 2^(-zs^2) is one FMUL + MUFU.EX2;
// |zs| is clamped at 9.6 (|z| = 8, erfc ~ 1e-29) to keep the exponential a normal number.  The polynomial
// coefficients carry the factor 1/2 of the half-step; the Gaussian's constant 1/sqrt(pi) is left to the caller, who
// folds it into its own factors.
#define SVB_SQRT_LOG2E 1.2011224087864498f
SVB_HD void erf_step_raw(float zs, float &half_1p_erf, float &g0) {
    const float a = fmin2(fabsf(zs), 9.6f);
    const float t = frcp(1.0f + (0.3275911f / SVB_SQRT_LOG2E) * a);
    float p = 0.5f * 1.061405429f;
    p = p * t - 0.5f * 1.453152027f;
    p = p * t + 0.5f * 1.421413741f;
    p = p * t - 0.5f * 0.284496736f;
    p = p * t + 0.5f * 0.254829592f;
    g0 = fexp2(-a * a);
    const float half_erfc = (p * t) * g0;
    half_1p_erf = zs >= 0.0f ? 1.0f - half_erfc : half_erfc;
}
#define SVB_INV_SQRT_PI 0.5641895835477563f
#ifndef SVB_SUM_FORM
#define SVB_SUM_FORM 1            // tuning switch (scratch/build_variant.sh): 0 = per-element derivative terms everywhere
#endif

template <uint32_t F>
struct AslRest {
    static constexpr bool CASL = (F & SVBASL_F_CASL) != 0;
    static constexpr bool ATT = (F & SVBASL_F_INFERATT) != 0;
    static constexpr bool ART = (F & (SVBASL_F_INFERART | SVBASL_F_ARTONLY)) != 0;
    static constexpr bool ARTONLY = (F & SVBASL_F_ARTONLY) != 0;
    static constexpr bool INFWM = (F & SVBASL_F_INFERWM) != 0 && !ARTONLY;
    // The WM signal is added only under incwm (aslrest.py:327); inferwm WITHOUT incwm (possible when pvcorr is
    // not used, aslrest.py:103-105 sets both only under pvcorr) still creates the fwm / deltwm / t1wm parameters
    // (aslrest.py:197-211,225-229) - they then have no effect on the prediction and only see their priors.
    static constexpr bool INCWM = (F & SVBASL_F_INCWM) != 0 && !ARTONLY;
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
    static constexpr bool kRegHeavy = false;
    static constexpr bool kPairedSamples = true;      // voxel_step.h: both samples of a Philox call per loop trip
    static constexpr int PA = P > 0 ? P : 1;

    static constexpr int xf(int) { return SVBASL_XF_IDENTITY; }   // all Normal (aslrest.py:184-246)
    static constexpr int ix(int i) { return i < 0 ? 0 : i; }

    static constexpr int kMaxNB = 8;
    struct Vox {              // per-voxel constants
        float pvgm, pvwm;
        // exp(-t_b/T1app) of the batch's time points (register-resident batches with a fixed T1 only): the
        // per-element exponential exp(-(t_b - delta)/T1app) then factorises into eb[b] * exp(delta/T1app)
        float eb[kMaxNB], ebw[kMaxNB];
    };

    template <class Acc>
    static constexpr bool use_eb() { return Acc::NB > 0 && Acc::NB <= kMaxNB && !T1; }

    template <class Acc>
    static SVB_HD void bind_times(const DevModel &m, Vox &v, const Acc &acc) {
        if (use_eb<Acc>() && TISS) {
#pragma unroll
            for (int b = 0; b < (Acc::NB > 0 && Acc::NB <= kMaxNB ? Acc::NB : 1); ++b) {
                v.eb[b] = fexp2(acc.time(b) * m.gm.nk);
                if (INCWM) v.ebw[b] = fexp2(acc.time(b) * m.wm.nk);
            }
        }
    }

    // One tissue compartment's per-sample terms.  `k` is a copy of the constant-bank rates when T1 is fixed
    // (the compiler keeps reading the constant bank) and per-sample values when T1 is inferred.
    struct Tissue {
        TissueRates k;
        float delt, tdp;      // delta, fl(tau + delta)             (mask thresholds, aslrest.py:362-363)
        float A;              // CASL: 2*T1app*exp(-delta/t1b) (aslrest.py:371); PASL: 2*exp(r*delta)
        float AE;             // CASL with eb[]: A * exp(delta/T1app), so that A*E = AE * eb[b]
        float Ab;             // CASL: -A / t1b, the constant of dS_during/ddelta = (1/t1b - q) F E - A/t1b
        float pvf, pv;        // pv*f, pv
        float dqdt1;          // dq/dt1 = -1/t1^2
    };

    struct Sample {
        Tissue gm, wm;
        // arterial (aslrest.py:404-419).  The time loop produces RAW terms: for CASL the per-sample factor
        // kc = 2 exp(-deltblood/t1b) is applied after the loop (fbk = fblood kc on the prediction, scale_grads on the
        // derivative sums); the Gaussian's 1/sqrt(pi) sits in the dz factors.
        float fb, fbk, kc;
        float thr_out;            // lead-out begins at deltblood + tau/2                         (aslrest.py:411)
        float zin_a, zin_b;       // lead-in:  z sqrt(log2 e) = zin_a t + zin_b                   (aslrest.py:413-423)
        float zout_b;             // lead-out: z sqrt(log2 e) = -t/leadscale' + zout_b            (aslrest.py:412,422)
        float dzin_c, dzin_t;     // lead-in: dz/ddeltblood / sqrt(pi) = dzin_c + dzin_t t
    };

    static SVB_HD Vox load_vox(const DevModel &m, int64_t w) {
        Vox v;
        v.pvgm = m.pvgm ? m.pvgm[w] : m.pvgm_s;
        v.pvwm = m.pvwm ? m.pvwm[w] : m.pvwm_s;
        return v;
    }

    template <bool T1_SAMPLED>
    static SVB_HD void prep_tissue(const DevModel &m, Tissue &ts, const TissueRates &fixed, float fc_pc, float f,
                                   float delt, float t1, float pv) {
        ts.delt = delt;
        ts.tdp = m.tau + delt;
        ts.pv = pv;
        ts.pvf = pv * f;
        if (T1_SAMPLED) {
            const float it1 = frcp(t1);
            ts.k = tissue_rates(it1 + fc_pc, m.tau, m.inv_t1b, CASL);
            ts.dqdt1 = -it1 * it1;
        } else {
            ts.k = fixed;
            ts.dqdt1 = 0.0f;
        }
        // exponentials as 2^(x log2 e) with the factor folded into the rate on the host where the rate is an option
        ts.A = CASL ? ts.k.two_iq * fexp2(delt * m.nl2e_inv_t1b) : 2.0f * fexp(ts.k.r * delt);
        ts.AE = (CASL && !T1_SAMPLED) ? ts.k.two_iq * fexp2(delt * ts.k.l2e_q_b) : 0.0f;
        ts.Ab = CASL ? -ts.A * m.inv_t1b : 0.0f;
    }

    // x[P]: model-space parameter values for this sample
    static SVB_HD Sample prep_sample(const DevModel &m, const Vox &v, const float *x) {
        Sample s;
        if (TISS) {
            const float delt = ATT ? x[ix(I_DELT)] : m.att;
            prep_tissue<T1>(m, s.gm, m.gm, m.fc_pc, x[ix(I_FTISS)], delt, T1 ? x[ix(I_T1)] : 0.0f, v.pvgm);
            if (INCWM) {
                const float fwm = INFWM ? x[ix(I_FWM)] : m.fwm;
                const float dwm = (I_DELTWM >= 0) ? x[ix(I_DELTWM)] : m.attwm;
                prep_tissue<(I_T1WM >= 0)>(m, s.wm, m.wm, m.fc_pc_wm, fwm, dwm, (I_T1WM >= 0) ? x[ix(I_T1WM)] : 0.0f,
                                           v.pvwm);
            }
        }
        if (ART) {
            s.fb = x[ix(I_FBLOOD)];
            const float db = (I_DELTBLOOD >= 0) ? x[ix(I_DELTBLOOD)] : m.artt;
            s.kc = CASL ? 2.0f * fexp2(db * m.nl2e_inv_t1b) : 1.0f;
            s.fbk = CASL ? s.fb * s.kc : s.fb;
            s.thr_out = db + m.half_tau;
            const float ls = fmin2(db, m.leadscale);
            const bool leadin_ok = ls > 0.0f;                                     // aslrest.py:419
            // tf.minimum routes the gradient to deltblood when it is the smaller argument: z_in = t/db - 1
            const bool own = db <= m.leadscale;
            const float ils = own ? frcp(leadin_ok ? ls : 1.0f) : m.inv_leadscale;
            // Without a lead-in (leadscale or deltblood <= 0) the signal before the lead-out is zero: z is pinned at
            // the clamp of the error-function step (1/2 erfc(8) ~ 1e-29) and its derivative terms vanish, which
            // spares the time loop a select per element.
            s.zin_a = leadin_ok ? ils * SVB_SQRT_LOG2E : 0.0f;
            s.zin_b = leadin_ok ? -db * s.zin_a : -9.6f;                          // z_in = (t - deltblood) / ls
            s.zout_b = db * m.inv_leadscale_s + m.tau_inv_leadscale_s;            // z_out = (tau + deltblood - t) / leadscale
            s.dzin_c = (own || !leadin_ok) ? 0.0f : -ils * SVB_INV_SQRT_PI;
            s.dzin_t = (own && leadin_ok) ? -ils * ils * SVB_INV_SQRT_PI : 0.0f;
        }
        return s;
    }

    // value S (per unit pv*f) and derivatives wrt delta and q of one tissue compartment at time t
    // USE_EB: ebt = exp(-t/T1app) is supplied (Vox::eb), saving the per-element exponential
    template <bool WANT_Q, bool USE_EB>
    static SVB_HD void tissue_eval(const DevModel &m, const Tissue &ts, float t, float ebt, float &S, float &dSdd,
                                   float &dSdq) {
        // the bolus has arrived / has passed (aslrest.py:362-363); tau >= 0, so `post` implies `arrived` and the
        // three-way masks become two nested selects
        const bool post = t > ts.tdp;
        const bool arrived = t > ts.delt;
        const float u = t - ts.delt;
        if (CASL) {
            // A * exp(-(t-delta)/T1app)
            const float FE = USE_EB ? ts.AE * ebt : ts.A * fexp2(u * ts.k.nk);
            const float Sd = ts.A - FE;                    // aslrest.py:372
            const float Sp = FE * ts.k.c1;                 // aslrest.py:373 with exp(tau q) folded into c1
            const float dd = FE * (m.inv_t1b - ts.k.q) + ts.Ab;          // -Sd/t1b - F E q with Sd = A - F E
            const float dp = FE * (ts.k.c1 * (ts.k.q - m.inv_t1b));
            S = post ? Sp : (arrived ? Sd : 0.0f);
            dSdd = post ? dp : (arrived ? dd : 0.0f);
            if (WANT_Q) {
                const float qd = FE * u - Sd * ts.k.iq;
                const float qp = Sp * (ts.k.tc1 - u - ts.k.iq);
                dSdq = post ? qp : (arrived ? qd : 0.0f);
            }
        } else {
            // factor*(exp(r t) - exp(r delt)) = 2 exp(-t q) exp(r delt) * (exp(r u) - 1)/r, which (unlike the
            // reference's float32 form) stays accurate when T1app is close to t1b (r -> 0)
            const float Be2 = (USE_EB ? ebt : fexp2(t * ts.k.nk)) * ts.A;   // 2 exp(-t/T1app) exp(r delt)
            const float e1u = em1r(ts.k.r, u);
            const float Sd = Be2 * e1u;                    // aslrest.py:379
            const float Sp = Be2 * ts.k.e1tau;             // aslrest.py:380
            S = post ? Sp : (arrived ? Sd : 0.0f);
            dSdd = post ? ts.k.r * Sp : (arrived ? -Be2 : 0.0f);
            if (WANT_Q) {
                const float qd = Be2 * dem1r(ts.k.r, u, e1u) - u * Sd;
                const float qp = Be2 * ts.k.de1tau - u * Sp;
                dSdq = post ? qp : (arrived ? qd : 0.0f);
            }
        }
    }

    // prediction at time t and the RAW derivative terms: d[] lacks the per-sample factors (pv, pv*f, fblood,
    // dq/dt1) that scale_grads() applies once per sample after the time loop
    template <bool USE_EB>
    static SVB_HD void eval(const DevModel &m, const Vox &v, const Sample &s, float t, int b, float &pred, float *d) {
        pred = 0.0f;
        if (TISS) {
            float S, dd, dq = 0.0f;
            tissue_eval<T1, USE_EB>(m, s.gm, t, USE_EB ? v.eb[USE_EB ? b : 0] : 0.0f, S, dd, dq);
            pred = s.gm.pvf * S;
            d[ix(I_FTISS)] = S;
            if (ATT) d[ix(I_DELT)] = dd;
            if (T1) d[ix(I_T1)] = dq;
            if (INCWM) {
                float Sw, ddw, dqw = 0.0f;
                tissue_eval<(I_T1WM >= 0), USE_EB>(m, s.wm, t, USE_EB ? v.ebw[USE_EB ? b : 0] : 0.0f, Sw, ddw, dqw);
                pred += s.wm.pvf * Sw;
                if (INFWM) d[ix(I_FWM)] = Sw;
                if (I_DELTWM >= 0) d[ix(I_DELTWM)] = ddw;
                if (I_T1WM >= 0) d[ix(I_T1WM)] = dqw;
            } else {
                if (INFWM) d[ix(I_FWM)] = 0.0f;               // parameters without a signal term
                if (I_DELTWM >= 0) d[ix(I_DELTWM)] = 0.0f;
                if (I_T1WM >= 0) d[ix(I_T1WM)] = 0.0f;
            }
        } else {
            if (T1) d[ix(I_T1)] = 0.0f;                    // artonly + infert1: parameter exists, unused
            if (I_T1WM >= 0) d[ix(I_T1WM)] = 0.0f;
        }
        if (ART) {
            const bool leadout = t > s.thr_out;                             // aslrest.py:411
            // z of aslrest.py:422-423 times sqrt(log2 e) (erf_step_raw): affine in t on either side of the switch
            const float za = leadout ? -m.inv_leadscale_s : s.zin_a;
            const float zb = leadout ? s.zout_b : s.zin_b;
            const float dz = leadout ? m.inv_leadscale_pi : (s.dzin_c + s.dzin_t * t);     // dz/ddeltblood / sqrt(pi)
            float h, g0;
            erf_step_raw(za * t + zb, h, g0);
            float A, dA;
            if (CASL) {                                                     // kc = 2 exp(-deltblood/t1b): per sample
                A = h;
                dA = g0 * dz - m.inv_t1b * h;
            } else {                                                        // kc = 2 exp(-t/t1b)   (aslrest.py:404-407)
                const float kc = 2.0f * fexp(-t * m.inv_t1b);
                A = kc * h;
                dA = kc * (g0 * dz);
            }
            pred += s.fbk * A;
            d[ix(I_FBLOOD)] = A;
            if (I_DELTBLOOD >= 0) d[ix(I_DELTBLOOD)] = dA;
        }
    }

    // per-sample factors of the raw derivative sums G[p] = sum_b r_b d_b[p]
    static SVB_HD void scale_grads(const Sample &s, float *G) {
        if (TISS) {
            G[ix(I_FTISS)] *= s.gm.pv;
            if (ATT) G[ix(I_DELT)] *= s.gm.pvf;
            if (T1) G[ix(I_T1)] *= s.gm.pvf * s.gm.dqdt1;
            if (INFWM && INCWM) G[ix(I_FWM)] *= s.wm.pv;
            if (I_DELTWM >= 0 && INCWM) G[ix(I_DELTWM)] *= s.wm.pvf;
            if (I_T1WM >= 0 && INCWM) G[ix(I_T1WM)] *= s.wm.pvf * s.wm.dqdt1;
        }
        if (ART) {
            if (CASL) G[ix(I_FBLOOD)] *= s.kc;
            if (I_DELTBLOOD >= 0) G[ix(I_DELTBLOOD)] *= s.fbk;
        }
    }

    // forward value only (Model.evaluate)
    static SVB_HD float predict(const DevModel &m, const Vox &v, const float *x, float t) {
        Sample s = prep_sample(m, v, x);
        float pred, d[PA];
        eval<false>(m, v, s, t, 0, pred, d);
        return pred;
    }

    // ---- CASL with a fixed T1: derivative sums without per-element derivative terms -----------------------------------
    // For a tissue compartment  dS/ddelta = -(1/t1b - q) S - [delta < t <= delta + tau] A q   (before arrival both
    // vanish; during the bolus S = A - F E and dS/ddelta = (1/t1b - q) F E - A/t1b; afterwards S = c1 F E and
    // dS/ddelta = (q - 1/t1b) S), so  sum_b r_b dS_b/ddelta = -(1/t1b - q) G_f - A q (sum_{arrived} r_b - sum_{post} r_b)
    // with G_f = sum_b r_b S_b, which the loop forms anyway; likewise the arterial term's  -h/t1b  part of
    // dA/ddeltblood is  -G_fblood / t1b.  The time loop then only carries values, two predicated adds and the products
    // with the residual.
    static constexpr bool kSumForm = SVB_SUM_FORM && CASL && !T1;

    struct TissueSums {
        float f, ra, rp;          // sum_b r_b S_b; sums of r_b over the arrived / the post-bolus time points
    };

    template <bool USE_EB>
    static SVB_HD float tissue_value(const Tissue &ts, float t, float ebt, bool &arrived, bool &post) {
        post = t > ts.tdp;
        arrived = t > ts.delt;
        const float FE = USE_EB ? ts.AE * ebt : ts.A * fexp2((t - ts.delt) * ts.k.nk);
        const float Sd = ts.A - FE;                        // aslrest.py:372
        const float Sp = FE * ts.k.c1;                     // aslrest.py:373
        return post ? Sp : (arrived ? Sd : 0.0f);
    }

    static SVB_HD void tissue_sums_close(const DevModel &m, const Tissue &ts, const TissueSums &u, float &G_f, float &G_d) {
        G_f = u.f;
        G_d = (ts.k.q - m.inv_t1b) * u.f - (ts.A * ts.k.q) * (u.ra - u.rp);
    }

    template <bool USE_EB, class Acc>
    static SVB_HD void run_sum_form(const DevModel &m, const Vox &v, const Sample &s, Acc &acc, int b, float t,
                                    TissueSums &ug, TissueSums &uw, float &G_fb, float &G_db) {
        float pred = 0.0f, Sg = 0.0f, Sw = 0.0f, h = 0.0f, gdz = 0.0f;
        bool arr_g = false, post_g = false, arr_w = false, post_w = false;
        if (TISS) {
            Sg = tissue_value<USE_EB>(s.gm, t, USE_EB ? v.eb[USE_EB ? b : 0] : 0.0f, arr_g, post_g);
            pred = s.gm.pvf * Sg;
            if (INCWM) {
                Sw = tissue_value<USE_EB>(s.wm, t, USE_EB ? v.ebw[USE_EB ? b : 0] : 0.0f, arr_w, post_w);
                pred += s.wm.pvf * Sw;
            }
        }
        if (ART) {
            const bool leadout = t > s.thr_out;                             // aslrest.py:411
            const float za = leadout ? -m.inv_leadscale_s : s.zin_a;
            const float zb = leadout ? s.zout_b : s.zin_b;
            const float dz = leadout ? m.inv_leadscale_pi : (s.dzin_c + s.dzin_t * t);
            float g0;
            erf_step_raw(za * t + zb, h, g0);
            gdz = g0 * dz;
            pred += s.fbk * h;
        }
        const float r = acc.resid(b, pred);
        if (TISS) {
            ug.f += r * Sg;
            if (ATT) {
                if (arr_g) ug.ra += r;
                if (post_g) ug.rp += r;
            }
            if (INCWM && INFWM) {
                uw.f += r * Sw;
                if (ATT) {
                    if (arr_w) uw.ra += r;
                    if (post_w) uw.rp += r;
                }
            }
        }
        if (ART) {
            G_fb += r * h;
            if (I_DELTBLOOD >= 0) G_db += r * gdz;
        }
    }

    // Visit every time point of the batch.  Acc supplies: static NB (compile-time batch size, 0 = dynamic),
    // n(), time(b), add(b, pred, d) / resid(b, pred) and the accumulated G[].
    template <class Acc>
    static SVB_HD void run(const DevModel &m, const Vox &v, const float *x, Acc &acc) {
        Sample s = prep_sample(m, v, x);
        if (kSumForm) {
            TissueSums ug = {0.0f, 0.0f, 0.0f}, uw = {0.0f, 0.0f, 0.0f};
            float G_fb = 0.0f, G_db = 0.0f;
            if (Acc::NB > 0) {
#pragma unroll
                for (int b = 0; b < (Acc::NB > 0 ? Acc::NB : 1); ++b)
                    run_sum_form<use_eb<Acc>()>(m, v, s, acc, b, acc.time(b), ug, uw, G_fb, G_db);
            } else {
                const int nb = acc.n();
                for (int b = 0; b < nb; ++b) run_sum_form<false>(m, v, s, acc, b, acc.time(b), ug, uw, G_fb, G_db);
            }
            float *G = acc.G;
#pragma unroll
            for (int p = 0; p < P; ++p) G[p] = 0.0f;             // parameters without a signal term stay at zero
            if (TISS) {
                float gf, gd;
                tissue_sums_close(m, s.gm, ug, gf, gd);
                G[ix(I_FTISS)] = gf;
                if (ATT) G[ix(I_DELT)] = gd;
                if (INCWM && INFWM) {
                    tissue_sums_close(m, s.wm, uw, gf, gd);
                    G[ix(I_FWM)] = gf;
                    if (I_DELTWM >= 0) G[ix(I_DELTWM)] = gd;
                }
            }
            if (ART) {
                G[ix(I_FBLOOD)] = G_fb;
                if (I_DELTBLOOD >= 0) G[ix(I_DELTBLOOD)] = G_db - m.inv_t1b * G_fb;
            }
        } else if (Acc::NB > 0) {
#pragma unroll
            for (int b = 0; b < (Acc::NB > 0 ? Acc::NB : 1); ++b) {
                float pred, d[PA];
                eval<use_eb<Acc>()>(m, v, s, acc.time(b), b, pred, d);
                acc.add(b, pred, d);
            }
        } else {
            const int nb = acc.n();
            for (int b = 0; b < nb; ++b) {
                float pred, d[PA];
                eval<false>(m, v, s, acc.time(b), b, pred, d);
                acc.add(b, pred, d);
            }
        }
        scale_grads(s, acc.G);
    }
};

}  // namespace svb
