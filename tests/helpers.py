"""
Shared machinery for the parity tests: builds the C-ABI descriptors from the oracle's configuration objects
and runs the same call either through libsvbasl.so on the GPU ("cuda") or through the host build of the
device headers ("hostsim", CPU-only check of the kernel arithmetic for the GPU-less container).
"""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import asl_models as om            # noqa: E402
from oracle import svb_engine as eng           # noqa: E402
from svb_models_asl_b200 import _lib as L      # noqa: E402

TIS = [2.05, 2.3, 2.55, 2.8, 3.05, 3.3]


def cfg_flags(cfg):
    f = 0
    f |= L.F_CASL if cfg.casl else 0
    f |= L.F_INFERATT if cfg.inferatt else 0
    f |= L.F_INFERART if cfg.inferart else 0
    f |= L.F_INCWM if cfg.incwm else 0
    f |= L.F_INFERWM if cfg.inferwm else 0
    f |= L.F_INFERT1 if cfg.infert1 else 0
    f |= L.F_ARTONLY if cfg.artonly else 0
    return f


class Backend:
    """Owns buffers on the right side (numpy for hostsim, torch.cuda for cuda) and calls the library."""

    def __init__(self, kind, defines=()):
        """defines: hostsim only - a variant build of the device headers (tests/hostsim/build_hostsim.py)."""
        self.kind = kind
        self.keep = []
        if kind == "cuda":
            self.lib = L.load()
            self.dev = torch.device("cuda:0")
        else:
            from tests.hostsim.build_hostsim import build
            self.lib = C.CDLL(build(tuple(defines)))
            self.lib.hostsim_step.argtypes = [C.POINTER(L.Model), C.POINTER(L.Engine), C.POINTER(L.Adam), C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            self.lib.hostsim_evaluate.argtypes = [C.POINTER(L.Model), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                  C.c_int, C.c_int, C.c_int64]
            self.lib.hostsim_n_params.argtypes = [C.POINTER(L.Model)]
            self.lib.hostsim_fill_eps.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                                  C.c_uint64, C.c_int64]
            self.lib.hostsim_fill_eps.restype = None
            self.lib.hostsim_sample_spatial.argtypes = [C.POINTER(L.Engine), C.c_int64, C.c_int64, C.c_void_p]
            self.lib.hostsim_sample_spatial.restype = None

    # ---- buffers ----
    def put(self, arr, dtype=np.float32):
        if arr is None:
            return None
        a = np.ascontiguousarray(np.asarray(arr), dtype=dtype)
        if self.kind == "cuda":
            t = torch.from_numpy(a).to(self.dev)
        else:
            t = a.copy()
        self.keep.append(t)
        return t

    def zeros(self, shape, dtype=np.float32):
        return self.put(np.zeros(shape, dtype=dtype), dtype)

    @staticmethod
    def ptr(t):
        if t is None:
            return None
        if isinstance(t, np.ndarray):
            return t.ctypes.data
        return t.data_ptr()

    def get(self, t):
        if isinstance(t, np.ndarray):
            return t.copy()
        torch.cuda.synchronize()
        return t.cpu().numpy()

    # ---- descriptors ----
    def model_desc(self, cfg, n_vox=None):
        m = L.Model()
        if isinstance(cfg, dict):                       # aslnn: {"weights": [...], "biases": [...]}
            m.kind = L.MODEL_ASLNN
            parts = []
            for w, b in zip(cfg["weights"], cfg["biases"]):
                parts += [np.asarray(w, np.float32).reshape(-1), np.asarray(b, np.float32).reshape(-1)]
            packed = np.ascontiguousarray(np.concatenate(parts))
            assert packed.size == 151
            self.keep.append(packed)
            m.nn_weights = packed.ctypes.data
            if cfg.get("tensor_cores"):
                m.flags = L.F_NN_TC
            return m
        m.kind = L.MODEL_ASLREST_DISP if cfg.disp else L.MODEL_ASLREST
        m.flags = cfg_flags(cfg)
        if cfg.disp:
            m.flags |= (L.F_DISP_INFER if cfg.infer_disp_params else 0)
            m.flags |= (L.F_DISP_ASWRITTEN if cfg.disp_postbolus == "as_written" else 0)
            m.conv_dt, m.conv_tmax = cfg.conv_dt, cfg.conv_tmax
            m.conv_nt = 1 + int(cfg.conv_tmax / cfg.conv_dt)
            m.s_fixed, m.sp_fixed = cfg.s_fixed, cfg.sp_fixed
        m.tau, m.t1b = cfg.tau, cfg.t1b
        m.t1, m.pc, m.fcalib, m.att = (float(np.mean(x)) for x in (cfg.t1, cfg.pc, cfg.fcalib, cfg.att))
        m.t1wm, m.pcwm, m.fcalibwm, m.attwm, m.fwm = (float(np.mean(x)) for x in
                                                      (cfg.t1wm, cfg.pcwm, cfg.fcalibwm, cfg.attwm, cfg.fwm))
        m.artt = cfg.artt
        m.leadscale = cfg.leadscale
        for name in ("pvgm", "pvwm"):
            v = np.asarray(getattr(cfg, name), dtype=np.float32)
            if v.ndim == 0:
                setattr(m, name + "_s", float(v))
                setattr(m, name, None)
            else:
                setattr(m, name + "_s", 0.0)
                setattr(m, name, self.ptr(self.put(v)))
        return m

    def evaluate(self, cfg, params, t, n_samples):
        """params [P,W,S,1], t [W|1,1,B] (reference shapes) -> [W,S,B] float32"""
        params = np.asarray(params, dtype=np.float32)
        P, W, S = params.shape[0], params.shape[1], params.shape[2]
        t = np.asarray(t, dtype=np.float32)
        B = t.shape[-1]
        n_rows = W * S
        m = self.model_desc(cfg)
        d_par = self.put(params.reshape(P, n_rows)) if P else None
        d_t = self.put(t.reshape(-1, B))
        d_out = self.zeros((n_rows, B))
        if self.kind == "cuda":
            L.check(self.lib.svbasl_evaluate(C.byref(m), self.ptr(d_par), self.ptr(d_t), self.ptr(d_out), n_rows, S, B,
                                             t.reshape(-1, B).shape[0], None))
        else:
            rc = self.lib.hostsim_evaluate(C.byref(m), self.ptr(d_par), self.ptr(d_t), self.ptr(d_out), n_rows, S, B,
                                           t.reshape(-1, B).shape[0])
            assert rc == 0, rc
        return self.get(d_out).reshape(W, S, B)

    def engine_desc(self, spec, state, data, tpts, eps=None, *, n_batch=None, t_row0=0, t_row_stride=1, seed=1,
                    neighbours=None, log_ak=None, w_begin=0, n_vox=None, vox_offset=0, n_vox_global=None,
                    grad_scale=None, ti=None, zoff=None, state_out=False, next_samples=False):
        """state [n_state,ld], data/tpts [T,ld], eps [P',S,ld] numpy (float64 ok) -> (Engine, buffers dict)"""
        ld = state.shape[1]
        e = L.Engine()
        e.n_vox = ld - w_begin if n_vox is None else n_vox
        e.w_begin = w_begin
        e.ld = ld
        e.vox_offset = vox_offset
        e.n_vox_global = n_vox_global or ld
        e.n_par = spec.n_par
        e.n_samples = spec.n_samples
        e.n_batch = n_batch or data.shape[0]
        e.t_full = spec.t_full
        e.latent = L.LATENT_NUMERIC if spec.latent == "numeric" else L.LATENT_ANALYTIC
        e.cov_llt = 1 if spec.cov == "LLt" else 0
        for i in range(spec.n_par):
            e.prior_type[i] = L.PRIOR_CODES[spec.prior_type[i]]
            e.prior_mean[i] = float(np.mean(spec.prior_mean[i]))
            e.prior_var[i] = float(spec.prior_var[i])
        e.ard_phi_max = spec.ard_phi_max or 0.0
        e.latent_weight = spec.latent_weight
        e.grad_scale = (1.0 / e.n_vox_global) if grad_scale is None else grad_scale
        bufs = {"state": self.put(state), "data": self.put(data)}
        bufs["tpts"] = self.put(tpts) if tpts is not None else None
        bufs["ti"] = self.put(ti) if ti is not None else None
        bufs["zoff"] = self.put(zoff) if zoff is not None else None
        bufs["eps"] = self.put(eps) if eps is not None else None
        bufs["nbr"] = self.put(neighbours, np.int32) if neighbours is not None else None
        bufs["log_ak"] = self.put(log_ak) if log_ak is not None else None
        bufs["ak_grad"] = self.zeros((4,), np.float64) if log_ak is not None else None
        bufs["sp"] = self.zeros((len(log_ak), spec.n_samples, ld)) if log_ak is not None else None
        bufs["sp_next"] = self.zeros((len(log_ak), spec.n_samples, ld)) if (log_ak is not None and next_samples) else None
        bufs["state_out"] = self.put(state) if state_out else None
        e.state = self.ptr(bufs["state"])
        e.state_out = self.ptr(bufs["state_out"])
        e.data = self.ptr(bufs["data"])
        e.tpts = self.ptr(bufs["tpts"])
        e.ti = self.ptr(bufs["ti"])
        e.zoff = self.ptr(bufs["zoff"])
        e.t_row0, e.t_row_stride = t_row0, t_row_stride
        e.eps = self.ptr(bufs["eps"])
        e.seed = seed
        e.neighbours = self.ptr(bufs["nbr"])
        e.spatial_samples = self.ptr(bufs["sp"])
        e.spatial_samples_out = self.ptr(bufs["sp_next"])
        e.log_ak = self.ptr(bufs["log_ak"])
        e.ak_grad = self.ptr(bufs["ak_grad"])
        return e, bufs

    def sample_spatial(self, e, bufs, step=0):
        """Pre-pass of a step with spatial priors: fills bufs["sp"] (svbasl_sample_spatial)."""
        if self.kind == "cuda":
            L.check(self.lib.svbasl_sample_spatial(C.byref(e), e.ld, step, self.ptr(bufs["sp"]), None))
        else:
            self.lib.hostsim_sample_spatial(C.byref(e), e.ld, step, self.ptr(bufs["sp"]))
        return self.get(bufs["sp"])

    def elbo_grad(self, m, e, n_state, step=0, nbt=0):
        cost = self.zeros((e.ld,))
        grad = self.zeros((n_state, e.ld))
        csum = self.zeros((1,), np.float64)
        if self.kind == "cuda":
            L.check(self.lib.svbasl_elbo_grad(C.byref(m), C.byref(e), step, self.ptr(cost), self.ptr(grad),
                                              self.ptr(csum), None))
        else:
            rc = self.lib.hostsim_step(C.byref(m), C.byref(e), None, step, self.ptr(cost), self.ptr(grad),
                                       self.ptr(csum), e.ak_grad, nbt)
            assert rc == 0, rc
        return self.get(cost), self.get(grad), float(self.get(csum)[0])

    def adam_desc(self, n_state, ld, lr, n_total, step0=0, n_iters=1, n_batches=1, b1=0.9, b2=0.999, eps=1e-8,
                  m=None, v=None):
        ad = L.Adam()
        steps = np.arange(1, n_total + 1, dtype=np.float64)
        lr_t = lr * np.sqrt(1 - b2 ** steps) / (1 - b1 ** steps)
        bufs = {"m": self.put(m) if m is not None else self.zeros((n_state, ld)),
                "v": self.put(v) if v is not None else self.zeros((n_state, ld)),
                "lr_t": self.put(lr_t)}
        ad.m, ad.v, ad.lr_t = self.ptr(bufs["m"]), self.ptr(bufs["v"]), self.ptr(bufs["lr_t"])
        ad.beta1, ad.beta2, ad.epsilon = b1, b2, eps
        ad.step0, ad.n_iters, ad.n_batches = step0, n_iters, n_batches
        return ad, bufs

    def step(self, m, e, ad, nbt=0):
        csum = self.zeros((max(1, ad.n_iters),), np.float64)
        nanc = self.zeros((1,), np.int64)
        if self.kind == "cuda":
            L.check(self.lib.svbasl_step(C.byref(m), C.byref(e), C.byref(ad), self.ptr(csum), self.ptr(nanc), None))
        else:
            rc = self.lib.hostsim_step(C.byref(m), C.byref(e), C.byref(ad), 0, None, None, self.ptr(csum), e.ak_grad,
                                       nbt)
            assert rc == 0, rc
        return self.get(csum), int(self.get(nanc)[0])

    def fill_eps(self, n_par, n_samples, ld, seed, step, vox_offset=0):
        eps = self.zeros((n_par, n_samples, ld))
        if self.kind == "cuda":
            L.check(self.lib.svbasl_fill_eps(self.ptr(eps), ld, ld, vox_offset, n_par, n_samples, seed, step, None))
        else:
            self.lib.hostsim_fill_eps(self.ptr(eps), ld, ld, vox_offset, n_par, n_samples, seed, step)
        return self.get(eps)


# ------------------------------------------------------------------------------------------------
def aslrest_spec(cfg, *, n_samples=10, t_full=6, latent="numeric", cov="LtL", ard=True, mrf=()):
    """EngineSpec with the priors AslRestModel.__init__ sets up (aslrest.py:183-246) + svb's noise parameter."""
    names = cfg.param_names()
    pm, pv, pt = [], [], []
    for n in names:
        if n == "ftiss":
            pm.append(1.5); pv.append(1e6); pt.append("N")
        elif n == "fwm":
            pm.append(0.5); pv.append(1e6); pt.append("N")
        elif n == "delttiss":
            pm.append(float(np.mean(cfg.att))); pv.append(1.0); pt.append("N")
        elif n == "deltwm":
            pm.append(float(np.mean(cfg.attwm))); pv.append(1.0); pt.append("N")
        elif n == "t1":
            pm.append(float(np.mean(cfg.t1))); pv.append(0.01); pt.append("N")
        elif n == "t1wm":
            pm.append(float(np.mean(cfg.t1wm))); pv.append(0.01); pt.append("N")
        elif n == "fblood":
            pm.append(0.0); pv.append(1e6); pt.append("A" if ard else "N")
        elif n == "deltblood":
            pm.append(cfg.artt); pv.append(1.0); pt.append("N")
        elif n == "s":                                                 # LogNormal(7.4, 2) geometric
            pm.append(math.log(7.4)); pv.append(math.log(2.0)); pt.append("N")
        elif n == "sp":
            pm.append(math.log(0.74)); pv.append(math.log(2.0)); pt.append("N")
        else:
            raise KeyError(n)
    for i in mrf:
        pt[i] = "M"
    pm.append(0.0); pv.append(math.log(2e5)); pt.append("N")          # noise: LogNormal(1, 2e5) (Appendix B)
    xf = [1 if n in ("s", "sp") else 0 for n in names] + [1]
    return eng.EngineSpec("aslrest", cfg, xf=xf, prior_type=pt, prior_mean=pm, prior_var=pv,
                          n_samples=n_samples, t_full=t_full, latent=latent, cov=cov)


def nn_spec(weights, biases, *, n_samples=10, t_full=6, latent="numeric", att=1.3):
    """EngineSpec for aslnn: ftiss LogNormal(1.5; prior var 1e6), delttiss FoldedNormal(att, 1) (aslnn.py:73-81)."""
    cfg = {"weights": weights, "biases": biases}
    return eng.EngineSpec("aslnn", cfg, xf=[1, 2, 1], prior_type=["N", "N", "N"],
                          prior_mean=[math.log(1.5), att, 0.0], prior_var=[math.log(1e6), 1.0, math.log(2e5)],
                          n_samples=n_samples, t_full=t_full, latent=latent)


def synth_problem(cfg, spec, W, rng, *, repeats=1, noise_sd=1.0, slicedt=0.0452, random_state=True):
    """Synthetic multi-PLD data in the style of scripts/gen_test_data.py + a non-trivial posterior state.
    -> dict(data [T,W], tpts [T,W], state [n_state,W] float64, truth)"""
    names = cfg.param_names()
    truth = {}
    for n in names:
        if n in ("ftiss", "fwm"):
            truth[n] = rng.uniform(1.0, 20.0, W)
        elif n in ("delttiss", "deltwm"):
            truth[n] = rng.uniform(0.6, 2.5, W)
        elif n in ("t1", "t1wm"):
            truth[n] = rng.uniform(1.0, 1.5, W)
        elif n == "fblood":
            truth[n] = rng.uniform(0, 10, W) * (rng.uniform(size=W) < 0.3)
        elif n == "deltblood":
            truth[n] = np.maximum(truth.get("delttiss", rng.uniform(0.6, 2.5, W)) - 0.3, 0.05)
        elif n == "s":
            truth[n] = rng.uniform(3.0, 12.0, W)
        elif n == "sp":
            truth[n] = rng.uniform(0.3, 2.0, W)
    z = rng.integers(0, 24, W)
    tis = np.repeat(np.asarray(TIS), repeats)
    tpts = (tis[:, None] + (z * slicedt)[None, :]).astype(np.float32)          # [T,W]
    par = [torch.as_tensor(truth[n]).reshape(W, 1, 1) for n in names]
    clean = om.evaluate(cfg, par, torch.as_tensor(tpts.astype(np.float64)).T.unsqueeze(1))[:, 0, :].T.numpy()
    data = (clean + rng.normal(0, noise_sd, clean.shape)).astype(np.float32)
    n = spec.n_par
    if random_state:
        rows = []
        for i, nme in enumerate(names):
            if nme in ("s", "sp"):                                             # internal value is the log
                rows.append(np.log(truth[nme]) + rng.normal(0, 0.2, W))
            else:
                rows.append(truth[nme] + rng.normal(0, 0.3, W))
        rows.append(rng.normal(0.3, 0.3, W))                                   # log noise variance
        for i in range(n):                                                     # log variances
            tight = i < len(names) and names[i] in ("t1", "t1wm")              # keep sampled T1 well away from 0
            rows.append(rng.normal(-6.0 if tight else -2.0, 0.5, W))
        for _ in range(spec.n_offdiag):
            rows.append(rng.normal(0, 0.05, W))
        for _ in spec.ard_params:
            rows.append(rng.normal(-3.0, 1.0, W))
        state = np.stack(rows, 0).astype(np.float32).astype(np.float64)
    else:
        state = None
    return {"data": data, "tpts": tpts, "state": state, "truth": truth}


def oracle_cost_grad(spec, prob, eps, rows=None, hyper=None, neighbours=None, grad_scale=None):
    state = torch.as_tensor(prob["state"], dtype=torch.float64)
    data = torch.as_tensor(prob["data"].astype(np.float64))
    t = torch.as_tensor(prob["tpts"].astype(np.float64))
    if rows is not None:
        data, t = data[rows], t[rows]
    hyper = torch.zeros(0, dtype=torch.float64) if hyper is None else torch.as_tensor(hyper, dtype=torch.float64)
    nb = None if neighbours is None else torch.as_tensor(neighbours, dtype=torch.int64)
    cost, gs, gh, aux = eng.cost_and_grad(spec, state, hyper, data, t, torch.as_tensor(eps, dtype=torch.float64), nb,
                                          grad_scale=grad_scale)
    return cost.numpy(), gs.numpy(), gh.numpy(), aux


def rel_err(a, b):
    """max |a-b| relative to the largest magnitude of the reference b, per leading row"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-30)
    return np.abs(a - b).max(axis=-1, keepdims=True) / scale


def trajectory_error(st, ref):
    """Posterior state after a few Adam iterations against the oracle's: per-element error relative to the largest
    magnitude of its state row -> dict(q50, q90, q99, max, frac_above_1e4).

    Why quantiles and not a plain bound: the first Adam step is lr*sign(g) and the next ones lr*m/sqrt(v), i.e. the
    optimiser divides by the gradient's own magnitude.  A voxel whose gradient for one variable nearly cancels (sum over
    samples and time points) turns a float32-level absolute gradient error into an O(lr) difference of that variable.
    Any two float32 evaluations of the same graph (the reference's TensorFlow on two devices included) part ways on
    such isolated voxels; what is checked is that the bulk agrees to the gradient tolerance and the stragglers stay
    inside the 2*lr-per-iteration envelope."""
    st, ref = np.asarray(st, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-30)
    err = np.abs(st - ref) / scale
    q = np.quantile(err, [0.5, 0.9, 0.99])
    return {"q50": float(q[0]), "q90": float(q[1]), "q99": float(q[2]), "max": float(err.max()),
            "max_abs": float(np.abs(st - ref).max()), "frac_above_1e-4": float((err > 1e-4).mean())}
