"""
ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement, in PyTorch ops (fp64 by default, differentiable by autograd), of
the reference's ASL forward models.  Every function cites the reference lines
it restates.  It is pinned against the *reference source itself* executed on a
numpy TensorFlow stand-in (tests/golden/make_golden.py -> tests/golden/*.npz,
checked by tests/test_oracle_golden.py) and against SURVEY.md Appendix D.

Shapes: every per-voxel quantity is broadcastable against ``t``; the usual call
has params ``[W,S,1]`` and ``t`` ``[W,1,B]`` giving ``[W,S,B]``
(reference: aslrest.py:248-261).
"""
from dataclasses import dataclass, field
import math

import numpy as np
import scipy.special
import torch


@dataclass
class AslConfig:
    """Resolved options of AslRestModel.__init__ (aslrest.py:69-246) as plain values."""
    casl: bool = True
    tau: float = 1.8
    t1b: float = 1.65
    t1: object = 1.3          # scalar or per-voxel [W] (aslrest.py:153,175)
    pc: object = 0.9
    fcalib: object = 0.01
    pvgm: object = 1.0
    att: object = 1.3
    artt: float = None        # arterial arrival when not inferred (Appendix C5)
    # white matter (aslrest.py:42-47, 213-219)
    incwm: bool = False
    inferwm: bool = False
    fwm: object = 0.0
    attwm: object = 1.6
    t1wm: object = 1.1
    pcwm: object = 0.8
    fcalibwm: object = 0.003
    pvwm: object = 0.0
    inferatt: bool = True
    inferart: bool = False
    infert1: bool = False
    artonly: bool = False
    leadscale: float = 0.01   # aslrest.py:232
    # dispersion (aslrest_disp.py:24-43)
    disp: bool = False
    infer_disp_params: bool = True
    conv_dt: float = 0.1
    conv_tmax: float = 5.0
    disp_postbolus: str = "intended"   # "as_written" reproduces gamma2-gamma2 == 0 (aslrest_disp.py:108)
    s_fixed: float = 7.4
    sp_fixed: float = 0.74

    def __post_init__(self):
        if self.artonly:
            self.inferart = True
        if self.artt is None:
            self.artt = float(np.mean(np.asarray(self.att))) - 0.3   # aslrest.py:86-87

    def param_names(self):
        """Parameter order of aslrest.py:183-246 (+ aslrest_disp.py:32-38)."""
        names = []
        if not self.artonly:
            names.append("ftiss")
            if self.inferatt:
                names.append("delttiss")
            if self.inferwm:
                names.append("fwm")
                if self.inferatt:
                    names.append("deltwm")
        if self.infert1:
            names.append("t1")
            if self.inferwm:
                names.append("t1wm")
        if self.inferart:
            names.append("fblood")
            if self.inferatt:
                names.append("deltblood")
        if self.disp and self.infer_disp_params:
            names += ["s", "sp"]
        return names


def _vox(x, like):
    """Per-voxel constant (scalar or [W]) -> tensor broadcastable as [W,1,1...] against `like`.
    The reference stores these as float32 node arrays (aslrest.py:142,153-157,215-219), so the
    value is rounded to float32 first whatever precision the evaluation then runs in."""
    x = torch.as_tensor(np.asarray(x, dtype=np.float32), dtype=like.dtype)
    if x.ndim == 0:
        return x
    return x.reshape([-1] + [1] * (like.ndim - 1))


def t1_apparent(cfg, t1, pc, fcalib, like):
    """-> (T1app, 1-exp(-tau/T1app), 1/T1app-1/t1b): the per-voxel rate terms of aslrest.py:366,373,376.
    fcalib, pc (and t1 unless inferred) are float32 node arrays in the reference (aslrest.py:142,
    153-157), so these sub-expressions are evaluated in float32 there whatever the precision of t;
    reproduced here so that the goldens agree to rounding."""
    ratio = np.asarray(fcalib, dtype=np.float32) / np.asarray(pc, dtype=np.float32)
    if torch.is_tensor(t1):
        t1_app = 1 / (1 / t1 + _vox(ratio, like))
        return t1_app, 1 - torch.exp(-cfg.tau / t1_app), 1 / t1_app - 1 / cfg.t1b
    t1 = np.asarray(t1, dtype=np.float32)
    t1_app = np.float32(1) / (np.float32(1) / t1 + ratio)
    post_fac = np.float32(1) - np.exp(-np.float32(cfg.tau) / t1_app)
    r = np.float32(1) / t1_app - np.float32(1 / cfg.t1b)
    return _vox(t1_app, like), _vox(post_fac, like), _vox(r, like)


def tissue_signal(cfg, t, f, delt, rates, pv):
    """Buxton tissue curve; restates aslrest.py:362-391 (masks :362-363, CASL :371-373,
    PASL :376-380, composition :387-391); rates from t1_apparent()."""
    t1_app, post_fac, r = rates
    post = t > (cfg.tau + delt)
    during = (t > delt) & ~post
    if cfg.casl:
        factor = 2 * t1_app * torch.exp(-delt / cfg.t1b)
        s_during = factor * (1 - torch.exp(-(t - delt) / t1_app))
        s_post = factor * torch.exp(-(t - cfg.tau - delt) / t1_app) * post_fac
    else:
        factor = 2 * torch.exp(-t / t1_app) / r
        s_during = factor * (torch.exp(r * t) - torch.exp(r * delt))
        s_post = factor * (torch.exp(r * (delt + cfg.tau)) - torch.exp(r * delt))
    zero = torch.zeros((), dtype=t.dtype)
    sig = torch.where(during, s_during, zero)
    sig = torch.where(post, s_post, sig)
    return pv * f * sig


def art_signal(cfg, t, fblood, deltblood):
    """Arterial curve with erf-smoothed edges; restates aslrest.py:404-430."""
    deltblood = deltblood + torch.zeros_like(t)
    if cfg.casl:
        kc = 2 * torch.exp(-deltblood / cfg.t1b)
    else:
        kc = 2 * torch.exp(-t / cfg.t1b)
    leadout = t > (deltblood + cfg.tau / 2)
    ls = torch.minimum(deltblood, torch.as_tensor(cfg.leadscale, dtype=t.dtype))
    leadin = ~leadout & (ls > 0)
    # the unused branch may divide by ls<=0; guard only the *unselected* lanes so that
    # autograd does not propagate NaN through torch.where (TF has the same hazard)
    ls_safe = torch.where(leadin, ls, torch.ones_like(ls))
    s_in = kc * 0.5 * (1 + torch.erf((t - deltblood) / ls_safe))
    s_out = kc * 0.5 * (1 + torch.erf(-(t - deltblood - cfg.tau) / cfg.leadscale))
    zero = torch.zeros((), dtype=t.dtype)
    sig = torch.where(leadin, s_in, zero)
    sig = torch.where(leadout, s_out, sig)
    return fblood * sig


# --------------------------------------------------------------------------
# Dispersion (aslrest_disp.py)
# --------------------------------------------------------------------------
class _IGammaC(torch.autograd.Function):
    """Q(a,x) = regularised upper incomplete gamma (tf.math.igammac, aslrest_disp.py:104-105),
    differentiable in both arguments: dQ/dx analytic, dQ/da by fp64 central differences."""

    @staticmethod
    def forward(ctx, a, x):
        a64, x64 = a.detach().double(), x.detach().double()
        ctx.save_for_backward(a64, x64)
        ctx.out_dtype = x.dtype
        q = scipy.special.gammaincc(a64.numpy(), x64.numpy())
        return torch.as_tensor(q).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        a, x = ctx.saved_tensors
        an, xn = a.numpy(), x.numpy()
        h = 1e-6 * np.maximum(1.0, np.abs(an))
        dqa = (scipy.special.gammaincc(an + h, xn) - scipy.special.gammaincc(an - h, xn)) / (2 * h)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            dqx = -np.exp((an - 1) * np.log(np.where(xn > 0, xn, 1.0)) - xn - scipy.special.gammaln(an))
        dqx = np.where(xn > 0, dqx, np.where(an > 1, 0.0, np.where(an == 1, -1.0, -np.inf)))
        dqa = np.where(xn > 0, dqa, 0.0)
        ga = g * torch.as_tensor(dqa).to(ctx.out_dtype)
        gx = g * torch.as_tensor(dqx).to(ctx.out_dtype)
        # reduce broadcast dims
        return _unbroadcast(ga, a.shape), _unbroadcast(gx, x.shape)


def _unbroadcast(g, shape):
    while g.ndim > len(shape):
        g = g.sum(0)
    for i, s in enumerate(shape):
        if s == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


def igammac(a, x):
    a, x = torch.broadcast_tensors(torch.as_tensor(a, dtype=x.dtype), x)
    return _IGammaC.apply(a.contiguous(), x.contiguous())


def aif_gammadisp(cfg, t, delt, s, sp):
    """Gamma-dispersed AIF; restates aslrest_disp.py:85-110 (Appendix C2 decision for the
    post-bolus term: default gamma2-gamma1 i.e. Q(k, s(t-d-tau)) - Q(k, s(t-d)))."""
    sp = torch.clamp(sp, max=10.0)
    delt = delt + torch.zeros_like(t)
    pre = t < delt
    post = t > (delt + cfg.tau)
    during = ~pre & ~post
    if cfg.casl:
        kc0 = 2 * torch.exp(-delt / cfg.t1b)
    else:
        kc0 = 2 * torch.exp(-t / cfg.t1b)
    k = 1 + sp
    g1 = igammac(k, s * torch.clamp(t - delt, 0, 1e6))
    g2 = igammac(k, s * torch.clamp(t - delt - cfg.tau, 0, 1e6))
    zero = torch.zeros((), dtype=t.dtype)
    kc = torch.where(during, kc0 * (1 - g1), zero)
    if cfg.disp_postbolus == "as_written":
        kc = torch.where(post, kc0 * (g2 - g2), kc)
    else:
        kc = torch.where(post, kc0 * (g2 - g1), kc)
    return kc


def conv_causal(aif, resid, dt):
    """conv_tf (aslrest_disp.py:148-171): C[i] = dt * sum_{m<=i} aif[m] * resid[i-m].
    `resid` may be [NT] or broadcastable [..., NT] (Appendix C3)."""
    nt = aif.shape[-1]
    out = []
    for i in range(nt):
        idx = torch.arange(i, -1, -1)
        out.append((aif[..., : i + 1] * resid[..., idx]).sum(-1))
    return torch.stack(out, -1) * dt


def interp_grid(t, tmax, curve):
    """tfp batch_interp_regular_1d_grid on [0,tmax] with constant extension (aslrest_disp.py:63)."""
    n = curve.shape[-1]
    pos = torch.clamp(t / tmax * (n - 1), 0, n - 1)
    lo = torch.clamp(torch.floor(pos.detach()).long(), 0, n - 2)
    frac = pos - lo
    shape = torch.broadcast_shapes(lo.shape[:-1], curve.shape[:-1])
    lo_b = lo.expand(*shape, lo.shape[-1])
    c_b = curve.expand(*shape, n)
    y0 = torch.gather(c_b, -1, lo_b)
    y1 = torch.gather(c_b, -1, lo_b + 1)
    return y0 + frac * (y1 - y0)


def disp_grid(cfg):
    """aslrest_disp.py:41-43 (nt = 1 + int(tmax/dt); linspace(0, tmax, nt))."""
    nt = 1 + int(cfg.conv_tmax / cfg.conv_dt)
    return torch.linspace(0.0, cfg.conv_tmax, nt, dtype=torch.float64), nt


def tissue_signal_disp(cfg, t, f, delt, rates, pv, s, sp):
    """aslrest_disp.py:48-64 with Appendix C1/C3/C4 decisions (per-voxel residue, pv applied)."""
    grid, _nt = disp_grid(cfg)
    grid = grid.to(t.dtype)
    aif = aif_gammadisp(cfg, grid, delt, s, sp)                     # [W,S,NT]
    t1_app = rates[0]
    resid = torch.exp(-grid / t1_app)                               # [W,1,NT] or [NT]
    curve = conv_causal(aif, resid + torch.zeros_like(aif), cfg.conv_dt)
    return pv * f * interp_grid(t, cfg.conv_tmax, curve)


# --------------------------------------------------------------------------
# Composition (aslrest.py:248-340)
# --------------------------------------------------------------------------
def evaluate(cfg, params, t):
    """params: sequence of tensors in cfg.param_names() order -> signal broadcast(params, t)."""
    names = cfg.param_names()
    if len(params) != len(names):
        raise ValueError("Model set up to infer %i parameters; this many parameter arrays "
                         "must be supplied" % len(names))              # aslrest.py:263-266
    p = dict(zip(names, params))
    like = t
    t1app = t1_apparent(cfg, p.get("t1", cfg.t1), cfg.pc, cfg.fcalib, like)
    t1app_wm = t1_apparent(cfg, p.get("t1wm", cfg.t1wm), cfg.pcwm, cfg.fcalibwm, like)
    extra = ()
    if cfg.disp:
        if cfg.infer_disp_params:
            extra = (p["s"], p["sp"])
        else:
            extra = (torch.as_tensor(cfg.s_fixed, dtype=t.dtype), torch.as_tensor(cfg.sp_fixed, dtype=t.dtype))
    tissue = (lambda *a: tissue_signal_disp(cfg, *a, *extra)) if cfg.disp else (lambda *a: tissue_signal(cfg, *a))
    sig = torch.zeros((), dtype=t.dtype)
    if not cfg.artonly:
        delt = p.get("delttiss", _vox(cfg.att, like))
        sig = tissue(t, p["ftiss"], delt, t1app, _vox(cfg.pvgm, like))
        if cfg.incwm:                                                   # aslrest.py:327-331
            fwm = p.get("fwm", _vox(cfg.fwm, like))
            deltwm = p.get("deltwm", _vox(cfg.attwm, like))
            sig = sig + tissue(t, fwm, deltwm, t1app_wm, _vox(cfg.pvwm, like))
    if cfg.inferart:                                                    # aslrest.py:336-338
        deltblood = p.get("deltblood", torch.as_tensor(cfg.artt, dtype=t.dtype))
        if cfg.disp:
            sig = sig + p["fblood"] * aif_gammadisp(cfg, t, deltblood, *extra)   # aslrest_disp.py:66-67
        else:
            sig = sig + art_signal(cfg, t, p["fblood"], deltblood)
    return sig + torch.zeros(torch.broadcast_shapes(t.shape, *[q.shape for q in params]), dtype=t.dtype)


# --------------------------------------------------------------------------
# aslnn surrogate (aslnn.py:93-126, 229-260)
# --------------------------------------------------------------------------
def mlp(x, weights, biases):
    """tanh(x W0 + b0) -> tanh(. W1 + b1) -> . W2 + b2 (aslnn.py:238-260)."""
    h = x
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = h @ torch.as_tensor(w, dtype=x.dtype) + torch.as_tensor(b, dtype=x.dtype)
        if i < n - 1:
            h = torch.tanh(h)
    return h


def evaluate_nn(weights, biases, ftiss, delt, t):
    """signal = ftiss * MLP([t, delt]) (aslnn.py:115-126); ftiss/delt [W,S,1], t [W|1,1,B]."""
    shape = torch.broadcast_shapes(t.shape, delt.shape)
    x = torch.stack([t.expand(shape), delt.expand(shape)], dim=-1)
    return ftiss * mlp(x, weights, biases).squeeze(-1)


def tpts(tis, repeats, shape, mask_vol, slicedt=0.0):
    """aslrest.py:432-456: per-voxel time points, grouped by TI, + z*slicedt -> [W,T] float32."""
    X, Y, Z = shape
    base = np.repeat(np.asarray(tis, dtype=np.float64), repeats)
    t = np.zeros((X, Y, Z, base.size), dtype=np.float32)
    for z in range(Z):
        t[:, :, z, :] = np.asarray([ti + z * slicedt for ti in tis for _ in range(repeats)])
    return t[np.asarray(mask_vol) > 0].reshape(-1, base.size)
