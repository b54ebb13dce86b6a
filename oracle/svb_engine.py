"""
ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the inference step that the external ``svb`` package wraps
around the reference plugins: reparameterised sampling from the per-voxel MVN
posterior, the model forward pass, the Gaussian-noise log-likelihood, the latent
loss (analytic KL or sample-based) against N / ARD / spatial-MRF priors, and the
TensorFlow-form Adam update.

PARITY UNPINNED: the svb source is not in /root/reference and svb/TensorFlow are
not installable here, so nothing below can be checked against the reference's
own output.  It follows SURVEY.md Appendix B (recalled svb semantics) with every
convention behind a named switch in ``EngineSpec``; the model forward pass inside
it (oracle/asl_models.py) *is* pinned to the reference source.  Gradients come
from torch autograd, i.e. independently of the hand-derived backward pass in the
CUDA kernels.

Written op-for-op like the TF graph (each elementary op materialises a [W,S,B]
tensor) so that, in float32 with all host threads, it also serves as the CPU
baseline of bench.py (``cpu_baseline.kind = "port"``).

Device state layout mirrored here (SoA, voxel fastest):
  state[k, w]:  k in [0,P')        posterior mean          (P' = P model params + noise, noise last)
                k in [P',2P')      posterior log-variance
                k in [2P',2P'+NL)  strict lower triangle of the Cholesky factor, rows (1,0),(2,0),(2,1),...
                then one log-phi row per ARD parameter
  hyper[j]:     log ak for the j-th spatial ('M') parameter (global scalars)
  eps[p, s, w]: standard normal draws
"""
from dataclasses import dataclass, field
import math

import numpy as np
import torch

from . import asl_models as om

XF_IDENTITY, XF_EXP, XF_ABS = 0, 1, 2


@dataclass
class EngineSpec:
    model: str                      # "aslrest" | "aslrest_disp" | "aslnn"
    cfg: object                     # om.AslConfig (aslrest*) or dict(weights=, biases=) (aslnn)
    xf: list                        # transform code per internal parameter, noise (XF_EXP) last
    prior_type: list                # "N" | "A" | "M" per internal parameter
    prior_mean: list                # internal-space prior mean per parameter (scalar or [W])
    prior_var: list                 # internal-space prior variance per parameter (scalar)
    n_samples: int = 10             # sample_size (asl_example.py:31)
    t_full: int = 6                 # total number of time points T (likelihood scale T/B)
    latent: str = "numeric"         # "numeric" = force_num_latent_loss (asl_example.py:41) | "analytic"
    cov: str = "LtL"                # covariance convention used by the KL/entropy: chol^T chol (svb) | "LLt"
    ard_phi_max: float = 1e6        # svb clips phi = exp(log_phi) to [0, 1e6]
    latent_weight: float = 1.0

    @property
    def n_par(self):
        return len(self.xf)

    @property
    def n_offdiag(self):
        return self.n_par * (self.n_par - 1) // 2

    @property
    def ard_params(self):
        return [i for i, t in enumerate(self.prior_type) if t == "A"]

    @property
    def spatial_params(self):
        return [i for i, t in enumerate(self.prior_type) if t == "M"]

    @property
    def n_state(self):
        return 2 * self.n_par + self.n_offdiag + len(self.ard_params)


def tril_index(n):
    """Strict lower-triangle (row, col) pairs in state order."""
    return [(i, j) for i in range(n) for j in range(i)]


def unpack_state(spec, state):
    """state [n_state, W] -> mean [W,P'], logvar [W,P'], chol [W,P',P'], logphi [W,n_ard]"""
    n = spec.n_par
    mean = state[:n].T
    logvar = state[n:2 * n].T
    chol = torch.diag_embed(torch.exp(0.5 * logvar))
    idx = tril_index(n)
    if idx:
        rows = torch.tensor([i for i, _ in idx])
        cols = torch.tensor([j for _, j in idx])
        off = torch.zeros_like(chol)
        off[:, rows, cols] = state[2 * n:2 * n + len(idx)].T
        chol = chol + off
    logphi = state[2 * n + len(idx):].T
    return mean, logvar, chol, logphi


def transform(code, theta):
    if code == XF_EXP:
        return torch.exp(theta)
    if code == XF_ABS:
        return torch.abs(theta)
    return theta


def model_predict(spec, ext, t):
    """ext: list of P tensors [W,S,1]; t [W,1,B] -> [W,S,B]"""
    if spec.model == "aslnn":
        return om.evaluate_nn(spec.cfg["weights"], spec.cfg["biases"], ext[0], ext[1], t)
    return om.evaluate(spec.cfg, ext, t)


def voxel_cost(spec, state, hyper, data, t, eps, neighbours=None):
    """
    Per-voxel negative free energy (svb ``cost``) for one batch of B time points.

    state [n_state,W]; hyper [n_spatial]; data, t [B,W]; eps [P',S,W];
    neighbours [W,6] int64 (-1 = none) when a spatial prior is present.
    -> cost [W], dict of intermediates
    """
    n, S = spec.n_par, spec.n_samples
    mean, logvar, chol, logphi = unpack_state(spec, state)
    B = data.shape[0]
    e = eps.permute(2, 0, 1)                                        # [W,P',S]
    theta = mean.unsqueeze(-1) + chol @ e                           # [W,P',S]   (sample = mean + chol eps)
    # materialise every operand at the full [W,S,B] shape, as TF's tile/broadcast kernels do (and because
    # torch's CPU broadcasting of [W,S,1] x [W,1,B] operands is ~50x slower than dense element-wise ops)
    full = (state.shape[1], S, B)
    if getattr(spec.cfg, "disp", False):      # the dispersion model works on its own [W,S,NT] grid
        ext = [transform(spec.xf[p], theta[:, p, :]).unsqueeze(-1) for p in range(n - 1)]
        pred = model_predict(spec, ext, t.T.unsqueeze(1).contiguous())
    else:
        ext = [transform(spec.xf[p], theta[:, p, :]).unsqueeze(-1).expand(full).contiguous() for p in range(n - 1)]
        pred = model_predict(spec, ext, t.T.unsqueeze(1).expand(full).contiguous())   # [W,S,B]
    log_nv = theta[:, n - 1, :]                                     # noise is LogNormal: var = exp(theta_n)
    nv = torch.exp(log_nv)
    ssd = torch.square(data.T.unsqueeze(1).expand(full).contiguous() - pred).sum(-1)          # [W,S]
    scale = spec.t_full / B
    recon = (0.5 * (log_nv * spec.t_full + scale * ssd / nv)).mean(1)

    # prior variance per parameter, with ARD phi where applicable
    pvar, pmean = [], []
    ard = spec.ard_params
    for p in range(n):
        m = torch.as_tensor(np.asarray(spec.prior_mean[p]), dtype=state.dtype)
        pmean.append(m + torch.zeros(state.shape[1], dtype=state.dtype))
        if p in ard:
            phi = torch.exp(logphi[:, ard.index(p)])
            if spec.ard_phi_max is not None:
                phi = torch.clamp(phi, 0, spec.ard_phi_max)
            pvar.append(1 / phi)
        else:
            pvar.append(torch.full((state.shape[1],), float(spec.prior_var[p]), dtype=state.dtype))
    pvar = torch.stack(pvar, 1)                                     # [W,P']
    pmean = torch.stack(pmean, 1)
    spatial = spec.spatial_params

    if spec.latent == "analytic" and not spatial:
        cov = chol.transpose(1, 2) @ chol if spec.cov == "LtL" else chol @ chol.transpose(1, 2)
        tr = (torch.diagonal(cov, dim1=1, dim2=2) / pvar).sum(1)
        quad = (torch.square(mean - pmean) / pvar).sum(1)
        latent = 0.5 * (tr + quad - n + torch.log(pvar).sum(1) - logvar.sum(1))
    else:
        entropy = -0.5 * logvar.sum(1)                              # -1/2 log det(cov)
        mlp = torch.zeros_like(entropy)
        for p in range(n):
            if p in spatial:
                ak = torch.exp(hyper[spatial.index(p)])
                x = theta[:, p, :]                                  # [W,S]
                nb = neighbours
                xn = x[nb.clamp(min=0)]                             # [W,6,S]
                d2 = torch.square(x.unsqueeze(1) - xn) * (nb >= 0).unsqueeze(-1)
                mlp = mlp + (0.5 * hyper[spatial.index(p)] - 0.25 * ak * d2.sum(1)).mean(1)
            else:
                z = torch.square(theta[:, p, :] - pmean[:, p:p + 1]) / pvar[:, p:p + 1]
                mlp = mlp + (-0.5 * torch.log(pvar[:, p:p + 1]) - 0.5 * z).mean(1)
        latent = entropy - mlp
    cost = recon + spec.latent_weight * latent
    return cost, {"pred": pred, "recon": recon, "latent": latent, "theta": theta}


def cost_and_grad(spec, state, hyper, data, t, eps, neighbours=None, grad_scale=None):
    """-> per-voxel cost [W], d(grad_scale*sum cost)/d state [n_state,W], d/d hyper.
    grad_scale defaults to 1/W (svb minimises the MEAN cost over voxels)."""
    state = state.detach().clone().requires_grad_(True)
    hyper = hyper.detach().clone().requires_grad_(True)
    cost, aux = voxel_cost(spec, state, hyper, data, t, eps, neighbours)
    gs = 1.0 / state.shape[1] if grad_scale is None else grad_scale
    total = cost.sum() * gs
    gstate, ghyper = torch.autograd.grad(total, [state, hyper], allow_unused=True)
    if ghyper is None:
        ghyper = torch.zeros_like(hyper)
    return cost.detach(), gstate, ghyper, aux


@dataclass
class Adam:
    """tf.train.AdamOptimizer update: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); x -= lr_t*m/(sqrt(v)+eps)
    (epsilon outside the bias correction, unlike torch.optim.Adam)."""
    lr: float = 0.01
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    step: int = 0
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)

    def update(self, named):
        """named: {key: (tensor, grad)} updated in place."""
        self.step += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.step) / (1 - self.b1 ** self.step)
        for k, (x, g) in named.items():
            if k not in self.m:
                self.m[k] = torch.zeros_like(x)
                self.v[k] = torch.zeros_like(x)
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            x.sub_(lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps))


def batch_rows(t_full, batch_size, i):
    """svb's non-sequential mini-batches over TIME POINTS: rows i, i+n_batches, ...."""
    n_batches = int(math.ceil(t_full / batch_size))
    return list(range(i % n_batches, t_full, n_batches)), n_batches


def initial_state(spec, post_mean, post_var, W, dtype=torch.float64):
    """post_mean/post_var: per internal parameter, scalar or [W] -> state [n_state,W]
    (mean, log var, zero off-diagonals, ARD log phi = log 1e-12)."""
    rows = []
    for m in post_mean:
        rows.append(torch.as_tensor(np.broadcast_to(np.asarray(m, dtype=np.float64), (W,)).copy(), dtype=dtype))
    for v in post_var:
        rows.append(torch.log(torch.as_tensor(np.broadcast_to(np.asarray(v, dtype=np.float64), (W,)).copy(),
                                              dtype=dtype)))
    rows += [torch.zeros(W, dtype=dtype)] * spec.n_offdiag
    rows += [torch.full((W,), math.log(1e-12), dtype=dtype)] * len(spec.ard_params)
    return torch.stack(rows, 0).contiguous()


def neighbour_table(coords, shape):
    """6-connected neighbours inside the mask: coords [W,3] int -> [W,6] int64, -1 where absent."""
    coords = np.asarray(coords)
    X, Y, Z = shape
    lut = -np.ones((X, Y, Z), dtype=np.int64)
    lut[coords[:, 0], coords[:, 1], coords[:, 2]] = np.arange(len(coords))
    out = -np.ones((len(coords), 6), dtype=np.int64)
    k = 0
    for axis in range(3):
        for step in (-1, 1):
            c = coords.copy()
            c[:, axis] += step
            ok = (c[:, axis] >= 0) & (c[:, axis] < shape[axis])
            cc = np.where(ok[:, None], c, 0)
            out[:, k] = np.where(ok, lut[cc[:, 0], cc[:, 1], cc[:, 2]], -1)
            k += 1
    return out


def fit(spec, state, hyper, data, t, n_iters, batch_size, lr, eps_fn, neighbours=None, history=None):
    """Reference training loop restated: Adam on the mean cost, strided time-point batches.
    data, t: [T,W]; eps_fn(iteration) -> eps [P',S,W].  Returns the final (state, hyper)."""
    opt = Adam(lr=lr)
    state = state.clone()
    hyper = hyper.clone()
    T = data.shape[0]
    for it in range(n_iters):
        rows, _nb = batch_rows(T, batch_size, it)
        cost, gs, gh, _ = cost_and_grad(spec, state, hyper, data[rows], t[rows], eps_fn(it), neighbours)
        named = {"state": (state, gs)}
        if hyper.numel():
            named["hyper"] = (hyper, gh)
        opt.update(named)
        if history is not None:
            history.append(float(cost.mean()))
    return state, hyper
