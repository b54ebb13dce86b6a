// model_disp.h - ASL kinetic model with gamma-kernel bolus dispersion: AIF(t) with regularised incomplete
// gamma functions, tissue curve = AIF convolved with the well-mixed residue exp(-t/T1app) on a regular grid,
// sampled at the time points by linear interpolation; arterial component = fblood * AIF(t).
//
// Takes over AslRestDisp (/root/reference/svb_models_asl/aslrest_disp.py): aif_gammadisp :69-110,
// resid_wellmix :133-146, conv_tf :148-171, the tfp interpolation :63, art_signal :66-67 - and the backward pass
// TensorFlow builds through tf.math.igammac (including d/da, which has no closed form).
// Decisions for the defects of the file as shipped: SURVEY.md Appendix C1-C4: the parent's
// 8-argument call is honoured, the residue is per voxel, pv is applied, and the post-bolus AIF is the intended
// kc*(Q(k,s(t-d-tau)) - Q(k,s(t-d))) unless SVBASL_F_DISP_ASWRITTEN asks for the shipped `gamma2 - gamma2` == 0.
//
// conv_tf is algebraically a causal discrete convolution times dt (golden-checked), and with an exponential
// residue that is the O(NT) recurrence C[i] = rho*C[i-1] + dt*AIF[i], rho = exp(-h/T1app);
// value and the derivatives wrt (delt, s, sp) are carried through it in forward mode, one sweep per sample.
#pragma once
#include "compat.h"
#include "dev_model.h"

namespace svb {

struct GammaConst {       // per (voxel, sample): everything about a = k that does not depend on x
    float a;              // k = 1 + min(sp, 10)
    float lg_a;           // lgamma(a)
    float psi_a;          // digamma(a)
    float ln_a, inv_a;
};

// digamma for a in (1, 11]: recurrence up to >= 6, then the asymptotic series
SVB_HD float digamma_f(float a) {
    float r = 0.0f;
#pragma unroll 1
    while (a < 6.0f) {
        r -= frcp(a);
        a += 1.0f;
    }
    const float i = frcp(a), i2 = i * i;
    return r + flog(a) - 0.5f * i - i2 * (1.0f / 12.0f - i2 * (1.0f / 120.0f - i2 * (1.0f / 252.0f)));
}

SVB_HD GammaConst gamma_const(float a) {
    GammaConst g;
    g.a = a;
    g.lg_a = lgammaf(a);
    g.psi_a = digamma_f(a);
    g.ln_a = flog(a);
    g.inv_a = frcp(a);
    return g;
}

// Q(a,x) = regularised upper incomplete gamma, with dQ/da and dQ/dx (forward-mode
// accumulation through the series / modified-Lentz continued fraction).  lnx = log(x).
SVB_HD void igammac_d(const GammaConst &g, float x, float lnx, float &Q, float &dQa, float &dQx) {
    const float a = g.a;
    if (!(x > 0.0f)) {
        Q = 1.0f; dQa = 0.0f; dQx = 0.0f;                 // a > 1: the density vanishes at 0
        return;
    }
    // density term x^(a-1) e^-x / Gamma(a) (= -dQ/dx) and the common prefactor exp(a ln x - x - lgamma(a))
    const float pre = fexp(a * lnx - x - g.lg_a);
    dQx = -pre * frcp(x);
    if (x < a + 1.0f) {
        // P = pre/a * sum_n t_n, t_0 = 1, t_n = t_{n-1} x/(a+n);  d t_n/da = -t_n H_n, H_n = sum_{k<=n} 1/(a+k)
        float t = 1.0f, sum = 1.0f, dsum = 0.0f, h = 0.0f, an = a;
#pragma unroll 1
        for (int n = 1; n < 200; ++n) {
            an += 1.0f;
            const float ian = frcp(an);
            t *= x * ian;
            h += ian;
            sum += t;
            dsum -= t * h;
            if (t < sum * 6e-8f) break;
        }
        const float d = pre * g.inv_a;                    // exp(a ln x - x - lgamma(a+1))
        const float dlog = lnx - (g.psi_a + g.inv_a);     // d log d / da = ln x - psi(a+1)
        Q = 1.0f - d * sum;
        dQa = -d * (dlog * sum + dsum);
    } else {
        // Q = pre * h, continued fraction b0 = x+1-a, a_i = -i(i-a), b_i = b_{i-1}+2 (modified Lentz),
        // carrying the a-derivatives of c, d and h
        const float tiny = 1e-30f;
        float b = x + 1.0f - a, db = -1.0f;
        float c = 1.0f / tiny, dc = 0.0f;
        float d = frcp(b), dd = -db * d * d;
        float h = d, dh = dd;
#pragma unroll 1
        for (int i = 1; i < 200; ++i) {
            const float fi = (float)i;
            const float an = -fi * (fi - a), dan = fi;
            b += 2.0f;
            float dn = an * d + b;
            float ddn = dan * d + an * dd + db;
            if (fabsf(dn) < tiny) dn = tiny;
            const float ic = frcp(c);
            float cn = b + an * ic;
            float dcn = db + (dan - an * dc * ic) * ic;
            if (fabsf(cn) < tiny) cn = tiny;
            d = frcp(dn);
            dd = -ddn * d * d;
            c = cn;
            dc = dcn;
            const float del = d * c;
            const float ddel = dd * c + d * dc;
            dh = dh * del + h * ddel;
            h *= del;
            if (fabsf(del - 1.0f) < 1.2e-7f) break;
        }
        Q = pre * h;
        dQa = pre * ((lnx - g.psi_a) * h + dh);
    }
}

// Q(a, x) along a (mostly) non-decreasing sequence of arguments with a fixed shape a - the situation of the AIF
// on the convolution grid, x_i = s (t_i - delt).  A fresh series / continued fraction per grid point costs up to
// ~60 data-dependent iterations and, worse, splits a warp three ways (series | continued fraction | lanes
// waiting): the first profile of this kernel ran with 8-12 of 32 lanes active (profiles/r1_notes.md section 6).
// Here the regularised lower function P and
//     J(a, x) = 1/Gamma(a) * int_0^x ln(t) t^(a-1) e^-t dt          (dP/da = J - psi(a) P)
// are obtained in two fixed-shape blocks:
//   1. arguments below 2 (and the restart of a sequence): the power series with a FIXED 14 terms - for x <= 2
//      and a > 1 the 14th term is < 2e-8 of the sum, so no convergence test and no lane-dependent trip count;
//   2. arguments above 2: advance from the previous argument (or from the series value at exactly 2) by 4-point
//      Gauss-Legendre quadrature of the density over pieces of width <= 1.  The integrand is analytic away from
//      t = 0 and the pieces start at t >= 2, where the quadrature error is ~1e-8 relative for every a in (1, 11]
//      (Bernstein-ellipse parameter rho >= 9.9, error ~ rho^-8) - below float32 rounding.
// Steps wider than 8 (s * h > 8: the kernel is narrower than one grid cell) fall back to igammac_d.
struct GammaRun {
    float x;          // last argument
    float P, J;
    bool live;
};

SVB_HD GammaRun gamma_run_start() {
    GammaRun r;
    r.x = 0.0f;
    r.P = 0.0f;
    r.J = 0.0f;
    r.live = false;
    return r;
}

// P(a,x), J(a,x) for 0 < x <= 2 by 14 terms of  P = x^a e^-x / Gamma(a+1) * sum_n x^n / ((a+1)...(a+n))
SVB_HD void gamma_series14(const GammaConst &g, float x, float lnx, float &P, float &J) {
    const float d = fexp(g.a * lnx - x - g.lg_a) * g.inv_a;
    float t = 1.0f, sum = 1.0f, dsum = 0.0f, hn = 0.0f, an = g.a;
#pragma unroll
    for (int n = 1; n <= 14; ++n) {
        an += 1.0f;
        const float ian = frcp(an);
        t *= x * ian;
        hn += ian;
        sum += t;
        dsum -= t * hn;
    }
    P = d * sum;
    const float dPa = d * ((lnx - (g.psi_a + g.inv_a)) * sum + dsum);
    J = dPa + g.psi_a * P;
}

SVB_HD void gamma_run_eval(const GammaConst &g, GammaRun &r, float x, float lnx, float &Q, float &dQa, float &dQx) {
    if (!(x > 0.0f)) {
        Q = 1.0f; dQa = 0.0f; dQx = 0.0f;
        return;
    }
    const float am1 = g.a - 1.0f;
    const bool cont = r.live && r.x >= 2.0f && x >= r.x;          // continue from the previous argument
    float from = cont ? r.x : fmin2(x, 2.0f);
    if (x - from > 8.0f) {
        igammac_d(g, x, lnx, Q, dQa, dQx);
        r.P = 1.0f - Q;
        r.J = g.psi_a * r.P - dQa;
    } else {
        if (!cont) gamma_series14(g, from, x < 2.0f ? lnx : 0.6931471805599453f, r.P, r.J);
        const float width = x - from;
        if (width > 0.0f) {
            const int n = (int)ceilf(width);                      // 1..8 pieces of width <= 1
            const float hw = 0.5f * width / (float)n;
            const float xi[2] = {0.3399810435848563f, 0.8611363115940526f};
            const float wt[2] = {0.6521451548625461f, 0.3478548451374538f};
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 1
            for (int k = 0; k < n; ++k) {
                const float mid = from + hw * (float)(2 * k + 1);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float o = hw * xi[j];
                    const float ta = mid - o, tb = mid + o;
                    const float la = flog(ta), lb = flog(tb);
                    const float ea = fexp(am1 * la - ta - g.lg_a), eb = fexp(am1 * lb - tb - g.lg_a);
                    s0 += wt[j] * (ea + eb);
                    s1 += wt[j] * (ea * la + eb * lb);
                }
            }
            r.P += hw * s0;
            r.J += hw * s1;
        }
        Q = 1.0f - r.P;
        dQa = g.psi_a * r.P - r.J;
        dQx = -fexp(am1 * lnx - x - g.lg_a);                     // -x^(a-1) e^-x / Gamma(a)
    }
    r.x = x;
    r.live = true;
}

// accumulator for a single time point (forward evaluation through the same sweep)
struct OnePointAcc {
    static constexpr int NB = 0;
    float t, out;
    SVB_HD int n() const { return 1; }
    SVB_HD float time(int) const { return t; }
    SVB_HD void add(int, float pred, const float *) { out = pred; }
};

template <uint32_t F>
struct AslDisp {
    static constexpr bool CASL = (F & SVBASL_F_CASL) != 0;
    static constexpr bool ATT = (F & SVBASL_F_INFERATT) != 0;
    static constexpr bool ART = (F & (SVBASL_F_INFERART | SVBASL_F_ARTONLY)) != 0;
    static constexpr bool ARTONLY = (F & SVBASL_F_ARTONLY) != 0;
    static constexpr bool TISS = !ARTONLY;
    static constexpr bool DISP = (F & SVBASL_F_DISP_INFER) != 0;

    // parameter order: aslrest.py:183-246 then s, sp (aslrest_disp.py:32-38)
    static constexpr int I_FTISS = TISS ? 0 : -1;
    static constexpr int I_DELT = (TISS && ATT) ? 1 : -1;
    static constexpr int N_A = TISS ? (1 + (ATT ? 1 : 0)) : 0;
    static constexpr int I_FBLOOD = ART ? N_A : -1;
    static constexpr int I_DELTBLOOD = (ART && ATT) ? N_A + 1 : -1;
    static constexpr int N_B = N_A + (ART ? (1 + (ATT ? 1 : 0)) : 0);
    static constexpr int I_S = DISP ? N_B : -1;
    static constexpr int I_SP = DISP ? N_B + 1 : -1;
    static constexpr int P = N_B + (DISP ? 2 : 0);
    static constexpr int PA = P > 0 ? P : 1;
    static constexpr bool kRegHeavy = true;

    static constexpr int xf(int p) { return (DISP && p >= N_B) ? SVBASL_XF_EXP : SVBASL_XF_IDENTITY; }
    static constexpr int ix(int i) { return i < 0 ? 0 : i; }

    struct Vox {
        float pvgm;
    };
    template <class Acc>
    static SVB_HD void bind_times(const DevModel &, Vox &, const Acc &) {}
    static SVB_HD Vox load_vox(const DevModel &m, int64_t w) {
        Vox v;
        v.pvgm = m.pvgm ? m.pvgm[w] : m.pvgm_s;
        return v;
    }

    struct Disp {             // per-sample dispersion terms
        float s, ln_s;
        bool sp_live;         // false when sp is clipped at 10 (zero gradient, aslrest_disp.py:85)
        GammaConst g;
    };

    static SVB_HD Disp prep_disp(const DevModel &m, const float *x) {
        Disp d;
        d.s = DISP ? x[ix(I_S)] : m.s_fixed;
        const float sp = DISP ? x[ix(I_SP)] : m.sp_fixed;
        d.sp_live = sp < 10.0f;
        d.ln_s = flog(d.s);
        d.g = gamma_const(1.0f + fmin2(sp, 10.0f));
        return d;
    }

    // AIF(t; delt) and its derivatives wrt delt, s, sp  (aslrest_disp.py:91-108)
    // r1 / r2: running incomplete-gamma states of the two argument sequences s (t - delt), s (t - delt - tau);
    // they persist across calls (gamma_run_eval restarts a sequence whose argument went down)
    static SVB_HD void aif(const DevModel &m, const Disp &dp, GammaRun &r1, GammaRun &r2, float t, float delt,
                           float kc_casl, float &A, float &dAd, float &dAs, float &dAsp) {
        A = dAd = dAs = dAsp = 0.0f;
        const float u = t - delt;
        if (u < 0.0f) return;                                   // pre-bolus (t < delt)
        const bool post = t > delt + m.tau;
        const float kc = CASL ? kc_casl : 2.0f * fexp(-t * m.inv_t1b);
        const float dkc = CASL ? -kc * m.inv_t1b : 0.0f;
        float q1, q1a, q1x;
        if (post && (m.flags & SVBASL_F_DISP_ASWRITTEN)) return;    // kc*(gamma2 - gamma2) == 0 as shipped
        gamma_run_eval(dp.g, r1, dp.s * u, dp.ln_s + flog(fmax2(u, 1e-30f)), q1, q1a, q1x);
        float q2 = 1.0f, q2a = 0.0f, q2x = 0.0f, u2 = 0.0f;
        if (post) {
            u2 = u - m.tau;
            gamma_run_eval(dp.g, r2, dp.s * u2, dp.ln_s + flog(fmax2(u2, 1e-30f)), q2, q2a, q2x);
        }
        // during: kc (1 - g1); post: kc (g2 - g1)
        const float base = q2 - q1;
        A = kc * base;
        // d/d delt: x_i = s (t - delt - ...) -> dx/d delt = -s
        dAd = dkc * base + kc * (-dp.s) * (q2x - q1x);
        dAs = kc * (q2x * u2 - q1x * u);
        dAsp = dp.sp_live ? kc * (q2a - q1a) : 0.0f;
    }

    template <class Acc>
    static SVB_HD void run(const DevModel &m, const Vox &v, const float *x, Acc &acc) {
        const Disp dp = prep_disp(m, x);
        const int nb = acc.n();
        const float fb = ART ? x[ix(I_FBLOOD)] : 0.0f;
        const float deltb = (I_DELTBLOOD >= 0) ? x[ix(I_DELTBLOOD)] : m.artt;
        const float kcb = CASL ? 2.0f * fexp(-deltb * m.inv_t1b) : 0.0f;
        // derivatives are with respect to the model-space values (s, sp); the LogNormal chain factor
        // d exp(theta)/d theta is applied by the engine (voxel_step.h)

        // arterial part at one time point, added to (pred, d)
        GammaRun rb1 = gamma_run_start(), rb2 = gamma_run_start();
        auto arterial = [&](float t, float &pred, float *d) {
            if (!ART) return;
            float A, dAd, dAs, dAsp;
            aif(m, dp, rb1, rb2, t, deltb, kcb, A, dAd, dAs, dAsp);
            pred += fb * A;
            d[ix(I_FBLOOD)] = A;
            if (I_DELTBLOOD >= 0) d[ix(I_DELTBLOOD)] = fb * dAd;
            if (DISP) {
                d[ix(I_S)] += fb * dAs;
                d[ix(I_SP)] += fb * dAsp;
            }
        };

        if (!TISS) {
            for (int b = 0; b < nb; ++b) {
                float pred = 0.0f, d[PA];
#pragma unroll
                for (int p = 0; p < P; ++p) d[p] = 0.0f;
                arterial(acc.time(b), pred, d);
                acc.add(b, pred, d);
            }
            return;
        }

        const float f = x[ix(I_FTISS)];
        const float delt = ATT ? x[ix(I_DELT)] : m.att;
        const float kct = CASL ? 2.0f * fexp(-delt * m.inv_t1b) : 0.0f;
        const float pvf = v.pvgm * f;
        const int nt = m.conv_nt;
        const float h = m.conv_h;                 // grid step of linspace(0, tmax, nt)  (aslrest_disp.py:43)
        const float to_pos = m.conv_inv_h;        // t -> grid position
        const float rho = m.conv_rho;             // exp(-h/T1app)

        // The sweep runs over k = i - i0, i0 = first grid point at or after bolus arrival (C == 0 before it), so
        // that all voxels of a warp enter the expensive first points of the incomplete-gamma sequences (series
        // phase) in the same loop iteration and then advance by quadrature together: with the loop over the
        // absolute grid index the arrival-time spread left ~9 of 32 lanes active (profiles/r1_notes.md section 6).
        // Time points are handled in chunks of kChunk: tissue curve at the chunk's time points first (kept in a
        // small local array), then the arterial term + residual accumulation in a loop that is uniform over lanes.
        constexpr int kChunk = 6;
        // i0 from delt/h, then corrected by at most one so that it satisfies aif()'s own float test u = t_i - delt
        // >= 0 exactly; clamped to [0, nt] first (a sample with a huge or non-finite delt sweeps nothing)
        const float p0 = delt * to_pos;
        int i0 = p0 > (float)nt ? nt : (p0 > 0.0f ? (int)ceilf(p0) : 0);
        if ((float)i0 * h - delt < 0.0f) ++i0;
        else if (i0 > 0 && (float)(i0 - 1) * h - delt >= 0.0f) --i0;
        for (int b0 = 0; b0 < nb; b0 += kChunk) {
            const int nbc = nb - b0 < kChunk ? nb - b0 : kChunk;
            int lo_b[kChunk];
            float tis[kChunk][4];                 // S, dS/d delt, dS/d s, dS/d sp at the chunk's time points
            int last = 0;
#pragma unroll 1
            for (int j = 0; j < nbc; ++j) {
                const float pos = fmin2(fmax2(acc.time(b0 + j) * to_pos, 0.0f), (float)(nt - 1));
                int lo = (int)pos;
                lo = lo > nt - 2 ? nt - 2 : lo;
                lo_b[j] = lo;
                last = lo + 1 > last ? lo + 1 : last;
                tis[j][0] = tis[j][1] = tis[j][2] = tis[j][3] = 0.0f;
            }
            float C = 0.0f, Cd = 0.0f, Cs = 0.0f, Csp = 0.0f;
            GammaRun rt1 = gamma_run_start(), rt2 = gamma_run_start();
            // one trip count for the whole warp (lanes past their own last grid point idle), so that the lanes meet
            // again at the top of every step
            const int n_steps = warp_max(last - i0 + 1);
#pragma unroll 1
            for (int k = 0; k < n_steps; ++k) {
                warp_converge();
                const int i = i0 + k;
                if (i <= last) {
                    const float ti = (float)i * h;
                    float A, dAd, dAs, dAsp;
                    aif(m, dp, rt1, rt2, ti, delt, kct, A, dAd, dAs, dAsp);
                    const float pC = C, pCd = Cd, pCs = Cs, pCsp = Csp;
                    C = rho * C + m.conv_dt * A;
                    Cd = rho * Cd + m.conv_dt * dAd;
                    Cs = rho * Cs + m.conv_dt * dAs;
                    Csp = rho * Csp + m.conv_dt * dAsp;
                    // every time point of the chunk that falls in [grid[i-1], grid[i]] (linear interpolation,
                    // constant extension); intervals that end before i0 keep S = 0
#pragma unroll 1
                    for (int j = 0; j < nbc; ++j) {
                        if (i == 0 || lo_b[j] != i - 1) continue;
                        const float pos = fmin2(fmax2(acc.time(b0 + j) * to_pos, 0.0f), (float)(nt - 1));
                        const float fr = pos - (float)lo_b[j];
                        tis[j][0] = pC + fr * (C - pC);
                        tis[j][1] = pCd + fr * (Cd - pCd);
                        tis[j][2] = pCs + fr * (Cs - pCs);
                        tis[j][3] = pCsp + fr * (Csp - pCsp);
                    }
                }
            }
#pragma unroll 1
            for (int j = 0; j < nbc; ++j) {
                const float S = tis[j][0];
                float pred = pvf * S, d[PA];
#pragma unroll
                for (int p = 0; p < P; ++p) d[p] = 0.0f;
                d[ix(I_FTISS)] = v.pvgm * S;
                if (ATT) d[ix(I_DELT)] = pvf * tis[j][1];
                if (DISP) {
                    d[ix(I_S)] = pvf * tis[j][2];
                    d[ix(I_SP)] = pvf * tis[j][3];
                }
                arterial(acc.time(b0 + j), pred, d);
                acc.add(b0 + j, pred, d);
            }
        }
    }

    static SVB_HD float predict(const DevModel &m, const Vox &v, const float *x, float t) {
        OnePointAcc one;
        one.t = t;
        one.out = 0.0f;
        run(m, v, x, one);
        return one.out;
    }
};

}  // namespace svb
