"""
Device-side operators of the ASL SVB hot path, as thin Python objects over the C ABI (include/svbasl.h).

PyTorch is used for buffer ownership, streams and (multi-GPU) torch.distributed only; every kernel is in
libsvbasl.so.  Nothing here computes on the CPU: without the library or without a CUDA device these calls
raise.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib as L


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.SvbAslError("no CUDA device: the ASL SVB kernels only run on the GPU (there is no CPU fallback)")
    return L.load()


def device_array(arr, device=None, dtype=torch.float32):
    device = device or torch.device("cuda", torch.cuda.current_device())
    if torch.is_tensor(arr):
        return arr.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device=device, dtype=dtype).contiguous()


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def evaluate_model(model, params, tpts):
    """Model.evaluate for the ASL plugins (aslrest.py:248-340): params list of [W,S,1] (or [P,W,S,1]),
    tpts [W,1,B] / [1,1,B] / [n,B] -> CUDA tensor [W,S,B]."""
    lib = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(params, (list, tuple)):
        cols = [device_array(p, dev) for p in params]
        shape = torch.broadcast_shapes(*[c.shape for c in cols]) if cols else (1, 1, 1)
        par = torch.stack([c.expand(shape) for c in cols], 0) if cols else torch.zeros((0,) + tuple(shape), device=dev)
    else:
        par = device_array(params, dev)
    t = device_array(tpts, dev)
    # normalise shapes to params [P, W, S] and t [Wt, B]
    if par.ndim == 4:
        par = par[..., 0]
    elif par.ndim == 2:
        par = par[..., None]
    P, W, S = par.shape
    if t.ndim == 3:
        t = t.reshape(t.shape[0], t.shape[-1])
    elif t.ndim == 1:
        t = t.reshape(1, -1)
    Wt, B = t.shape
    if Wt not in (1, W):
        raise ValueError("time points have %i rows but parameters have %i voxels" % (Wt, W))
    n_rows = W * S
    out = torch.empty((W, S, B), device=dev, dtype=torch.float32)
    m, keep = model.kernel_model(lambda a: device_array(a, dev))
    par = par.reshape(P, n_rows).contiguous()
    t = t.contiguous()
    L.check(lib.svbasl_evaluate(C.byref(m), par.data_ptr() if P else None, t.data_ptr(), out.data_ptr(), n_rows, S, B,
                                Wt, _stream_ptr()))
    del keep
    return out


def device_init_stats(data, tpts=None):
    """Per-voxel statistics of the data for the posterior initialisers (aslrest.py:461-520, svb's noise
    initialiser), computed on the device by svbasl_init_stats: data [T, ld] CUDA tensor, tpts [T, ld] or None ->
    dict(mean_t, max_t, var_t (population variance), t_at_max (time of the FIRST maximum, tf.argmax)), each [ld]."""
    lib = _require_cuda()
    T, ld = data.shape
    outs = {k: torch.empty(ld, device=data.device, dtype=torch.float32) for k in ("mean_t", "max_t", "var_t", "t_at_max")}
    if tpts is None:
        outs["t_at_max"].zero_()
    L.check(lib.svbasl_init_stats(data.data_ptr(), tpts.data_ptr() if tpts is not None else None, ld, ld, T,
                                  outs["mean_t"].data_ptr(), outs["max_t"].data_ptr(), outs["var_t"].data_ptr(),
                                  outs["t_at_max"].data_ptr(), _stream_ptr()))
    return outs


class InitData(np.ndarray):
    """The `data` argument handed to a parameter's post_init(param, t, data) callback: the [W, T] host array the
    reference's callbacks expect (aslrest.py:461-520), carrying `device_stats` - the per-voxel mean / max /
    variance / time-of-maximum already reduced on the GPU - so that this package's own initialisers need not
    reduce 10 M x 48 values on the host; a third-party callback sees a plain ndarray."""
    device_stats = None

    def __new__(cls, array, stats=None):
        obj = np.asarray(array).view(cls)
        obj.device_stats = stats
        return obj

    def __array_finalize__(self, obj):
        self.device_stats = getattr(obj, "device_stats", None) if obj is not None and getattr(obj, "shape", None) == self.shape else None


_NN_TILES = {}


def nn_evaluate_tc(model, params, tpts, want_hidden=False, check=True):
    """AslNNModel.evaluate with the 10x10 hidden layer on the tensor cores (tcgen05 / TMEM / TMA, csrc/nn_tc.cu).
    Same arguments and result as evaluate_model(); optionally also returns the hidden pre-activations [W,S,B,10]."""
    lib = _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(params, (list, tuple)):
        cols = [device_array(p, dev) for p in params]
        shape = torch.broadcast_shapes(*[c.shape for c in cols])
        par = torch.stack([c.expand(shape) for c in cols], 0)
    else:
        par = device_array(params, dev)
    t = device_array(tpts, dev)
    if par.ndim == 4:
        par = par[..., 0]
    elif par.ndim == 2:
        par = par[..., None]
    P, W, S = par.shape
    if P != 2:
        raise ValueError("aslnn takes 2 parameters (ftiss, delttiss)")
    if t.ndim == 3:
        t = t.reshape(t.shape[0], t.shape[-1])
    elif t.ndim == 1:
        t = t.reshape(1, -1)
    Wt, B = t.shape
    if Wt not in (1, W):
        raise ValueError("time points have %i rows but parameters have %i voxels" % (Wt, W))
    m, keep = model.kernel_model(None)
    key = (id(model), dev.index, keep[0].tobytes())
    if key not in _NN_TILES:
        host = np.zeros(512, dtype=np.float32)
        L.check(lib.svbasl_nn_pack_weights(C.byref(m), host.ctypes.data))
        _NN_TILES.clear()
        _NN_TILES[key] = device_array(host, dev)
    tile = _NN_TILES[key]
    n_rows = W * S
    par = par.reshape(2, n_rows).contiguous()
    t = t.contiguous()
    out = torch.empty((W, S, B), device=dev, dtype=torch.float32)
    hidden = torch.empty((W, S, B, 10), device=dev, dtype=torch.float32) if want_hidden else None
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    # tpts rows: the kernel indexes time rows by parameter row / (n_rows / n_t_rows); [W,B] tpts repeat over S samples
    L.check(lib.svbasl_nn_evaluate_tc(C.byref(m), tile.data_ptr(), par.data_ptr(), t.data_ptr(), out.data_ptr(),
                                      hidden.data_ptr() if want_hidden else None, n_rows, B, Wt, status.data_ptr(),
                                      _stream_ptr()))
    if check and int(status.item()) != 0:      # host sync; pass check=False inside timed loops
        raise L.SvbAslError("tensor-core MLP kernel: a bounded wait expired (TMEM/MMA pipeline did not complete)")
    return (out, hidden) if want_hidden else out


class _DevicePointer:
    """Raw device allocation exposed through __cuda_array_interface__ so torch can alias it (float32, C order)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class FusedSvb:
    """
    The per-iteration graph of svb's SvbFit for one shard of voxels, as one kernel launch:
    sample -> model -> log-likelihood -> latent loss -> gradients -> Adam.

    All arrays are SoA, voxel-fastest, row stride `ld`.  `state` rows: mean[P'], logvar[P'], off-diagonal
    Cholesky rows, ARD log-phi rows (include/svbasl.h).

    Spatial ("M") priors: the step reads the six neighbours' theta samples of this iteration from
    `sp_bufs[sp_cur]`; the samples of the NEXT iteration go into the other buffer - and, sharded over several GPUs,
    straight into the adjacent ranks' halo columns over NVLink peer memory - and the iteration ends with the all-reduce
    of d(cost)/d(log ak), the Adam step on log ak and the advance of the device-resident iteration counter.
    `halo_mode`: "peer" (peer-memory stores + all-reduce over peer-memory mailboxes), "peer+nccl" (peer stores, NCCL
    all-reduce), "nccl" (NCCL send/recv of the halo samples, NCCL all-reduce).
    """

    def __init__(self, model, data, tpts=None, *, ti=None, zoff=None, n_samples=10, batch_size=None,
                 latent="numeric", cov_llt=False, learning_rate=0.01, seed=1, prior_types=None, prior_means=None,
                 prior_vars=None, n_vox_global=None, vox_offset=0, halo=(0, 0), neighbours=None, ak_init=1e-5,
                 ard_phi_max=1e6, latent_weight=1.0, device=None, adam=(0.9, 0.999, 1e-8), max_steps=100000):
        self.lib = _require_cuda()
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.model = model
        self.mdesc, self._keep = model.kernel_model(lambda a: device_array(a, self.dev))
        self.data = device_array(data, self.dev)                       # [T, ld]
        self.T, self.ld = self.data.shape
        self.tpts = device_array(tpts, self.dev) if tpts is not None else None
        self.ti = device_array(ti, self.dev) if ti is not None else None
        self.zoff = device_array(zoff, self.dev) if zoff is not None else None
        self.halo = halo
        self.n_vox = self.ld - halo[0] - halo[1]
        self.n_vox_global = n_vox_global or self.n_vox
        self.vox_offset = vox_offset
        self.S = n_samples
        self.B = batch_size or self.T
        self.n_batches = int(math.ceil(self.T / self.B))
        if self.T % self.n_batches:
            raise ValueError("time points (%i) must divide into equal strided batches (batch_size=%i)" % (self.T, self.B))
        self.B = self.T // self.n_batches
        P = self.lib.svbasl_model_n_params(C.byref(self.mdesc))
        L.check(P)
        self.P, self.N = P, P + 1
        self.prior_types = list(prior_types)
        self.prior_means = [float(np.mean(v)) for v in prior_means]
        self.prior_vars = [float(v) for v in prior_vars]
        assert len(self.prior_types) == self.N
        self.ard = [i for i, t in enumerate(self.prior_types) if t == "A"]
        self.mrf = [i for i, t in enumerate(self.prior_types) if t == "M"]
        self.NL = self.N * (self.N - 1) // 2
        self.n_state = 2 * self.N + self.NL + len(self.ard)
        self.latent = L.LATENT_NUMERIC if (latent == "numeric" or self.mrf) else L.LATENT_ANALYTIC
        self.cov_llt = bool(cov_llt)
        self.lr = learning_rate
        self.b1, self.b2, self.adam_eps = adam
        self.seed = seed
        self.ard_phi_max = ard_phi_max or 0.0
        self.latent_weight = latent_weight
        self.step_count = 0
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=self.dev, dtype=dt)  # noqa: E731
        self.state = z(self.n_state, self.ld)
        self.m = z(self.n_state, self.ld)
        self.v = z(self.n_state, self.ld)
        steps = np.arange(1, max_steps + 1, dtype=np.float64)
        self.lr_t_host = (learning_rate * np.sqrt(1 - self.b2 ** steps) / (1 - self.b1 ** steps)).astype(np.float32)
        self.lr_t = device_array(self.lr_t_host, self.dev)
        self.max_fuse = 64
        self.cost_hist = z(max_steps + self.max_fuse, dt=torch.float64)   # summed cost of every iteration
        self.nan_count = z(1, dt=torch.int64)
        self.neighbours = device_array(neighbours, self.dev, torch.int32) if neighbours is not None else None
        self.eps = None
        # how a spatial iteration is launched (SVBASL_SPATIAL_FLOW; see _launch_spatial_iteration): "prepass" = step
        # kernel + next iteration's pre-pass kernel + hyper step; "separate_tail" = next samples drawn inside the step
        # kernel + hyper step; "fused" = everything in one launch
        self.spatial_flow = os.environ.get("SVBASL_SPATIAL_FLOW", "prepass")
        self.plan = None            # ShardPlan of a spatial prior sharded over several ranks
        self.halo_mode = None
        self.reduce_fn = None       # sums a small tensor over all ranks in place (NCCL / gloo)
        self.graphs = None          # enable_graph(): one CUDA-graph replay per spatial iteration
        self.peers = None
        self._shared = []
        if self.mrf:
            if self.neighbours is None:
                raise ValueError("spatial prior needs a neighbour table")
            n_sp = len(self.mrf)
            self.log_ak = torch.full((n_sp,), math.log(ak_init), device=self.dev, dtype=torch.float32)
            self.ak_m, self.ak_v = z(n_sp), z(n_sp)
            self.ak_grad = z(L.MAX_SPATIAL, dt=torch.float64)
            # neighbour-sample buffers [n_sp, S, ld], ping-pong: iteration t reads sp_bufs[sp_cur] and writes the
            # samples of t+1 into the other one
            self.sp_bufs = [z(n_sp, self.S, self.ld), z(n_sp, self.S, self.ld)]
            self.sp_cur, self.sp_valid = 0, False
            self.step_dev = z(1, dt=torch.int64)                          # iteration counter, device resident
            self.done_ctas = z(1, dt=torch.int32)
            self.peer_status = z(1, dt=torch.int32)
        else:
            self.log_ak = self.ak_grad = self.sp_bufs = None

    # ---- descriptors ----
    def engine_desc(self, row0=0, for_step=False):
        """`for_step`: descriptor of a fused spatial iteration (device counter, next-sample buffer, peer pointers)."""
        e = L.Engine()
        e.n_vox, e.w_begin, e.ld = self.n_vox, self.halo[0], self.ld
        e.vox_offset = self.vox_offset - self.halo[0]
        e.n_vox_global = self.n_vox_global
        e.n_par, e.n_samples, e.n_batch, e.t_full = self.N, self.S, self.B, self.T
        e.latent, e.cov_llt = self.latent, int(self.cov_llt)
        for i in range(self.N):
            e.prior_type[i] = L.PRIOR_CODES[self.prior_types[i]]
            e.prior_mean[i] = self.prior_means[i]
            e.prior_var[i] = self.prior_vars[i]
        e.ard_phi_max, e.latent_weight = self.ard_phi_max, self.latent_weight
        e.grad_scale = 1.0 / self.n_vox_global
        e.state = self.state.data_ptr()
        e.data = self.data.data_ptr()
        e.tpts = self.tpts.data_ptr() if self.tpts is not None else None
        e.ti = self.ti.data_ptr() if self.ti is not None else None
        e.zoff = self.zoff.data_ptr() if self.zoff is not None else None
        e.t_row0, e.t_row_stride = row0, self.n_batches
        e.eps = self.eps.data_ptr() if self.eps is not None else None
        e.seed = self.seed
        if self.mrf:
            e.neighbours = self.neighbours.data_ptr()
            e.spatial_samples = self.sp_bufs[self.sp_cur].data_ptr()
            e.log_ak = self.log_ak.data_ptr()
            e.ak_grad = self.ak_grad.data_ptr()
            if for_step:
                e.step_dev = self.step_dev.data_ptr()
                e.spatial_samples_out = self.sp_bufs[1 - self.sp_cur].data_ptr()
                if self.peers:
                    for side in ("lo", "hi"):
                        if side in self.peers:
                            p = self.peers[side]
                            # the neighbour's buffer that plays the "next samples" role in the same iteration
                            setattr(e, "peer_" + side, p["ptrs"][1 - self.sp_cur])
                            setattr(e, "peer_%s_ld" % side, p["ld"])
                            setattr(e, "peer_%s_shift" % side, p["shift"])
                            first, count = self._mirror[side]
                            setattr(e, "peer_%s_first" % side, first)
                            setattr(e, "peer_%s_count" % side, count)
        return e

    def adam_desc(self, n_iters=1):
        ad = L.Adam()
        ad.m, ad.v, ad.lr_t = self.m.data_ptr(), self.v.data_ptr(), self.lr_t.data_ptr()
        ad.beta1, ad.beta2, ad.epsilon = self.b1, self.b2, self.adam_eps
        ad.step0, ad.n_iters, ad.n_batches = self.step_count, n_iters, self.n_batches
        return ad

    def hyper_desc(self):
        """Fused tail of a spatial iteration (svbasl_hyper): single GPU (world = 1) or over the ranks' mailboxes."""
        h = L.Hyper()
        h.log_ak, h.m, h.v = self.log_ak.data_ptr(), self.ak_m.data_ptr(), self.ak_v.data_ptr()
        h.lr_t, h.step_dev, h.done_ctas = self.lr_t.data_ptr(), self.step_dev.data_ptr(), self.done_ctas.data_ptr()
        h.beta1, h.beta2, h.epsilon = self.b1, self.b2, self.adam_eps
        h.n_spatial = len(self.mrf)
        h.status = self.peer_status.data_ptr()
        if self.plan is not None and self.plan.world > 1:
            h.rank, h.world = self.plan.rank, self.plan.world
            for r in range(self.plan.world):
                h.mailboxes[r] = self._mail_ptrs[r]
        else:
            h.rank, h.world = 0, 1
        return h

    # ---- state ----
    def set_posterior(self, means, variances):
        """means/variances: per internal parameter (noise last), scalar or [n_vox] -> state rows"""
        sl = slice(self.halo[0], self.halo[0] + self.n_vox)
        self.state.zero_()
        for i, (mu, var) in enumerate(zip(means, variances)):
            self.state[i, sl] = device_array(np.broadcast_to(np.asarray(mu, dtype=np.float32), (self.n_vox,)).copy(),
                                             self.dev)
            self.state[self.N + i, sl] = torch.log(device_array(
                np.broadcast_to(np.asarray(var, dtype=np.float32), (self.n_vox,)).copy(), self.dev))
        for k in range(len(self.ard)):
            self.state[2 * self.N + self.NL + k] = math.log(1e-12)
        self.m.zero_()
        self.v.zero_()
        self.cost_hist.zero_()
        self.step_count = 0
        if self.mrf:
            if self.graphs is not None:
                raise ValueError("set_posterior() after enable_graph(): the captured launches are bound to a buffer parity")
            self.step_dev.zero_()
            self.sp_cur, self.sp_valid = 0, False

    # ---- the hot path ----
    def step(self, n_iters=1, want_cost=True):
        """`n_iters` fused iterations (ELBO + gradient + Adam) in ONE launch.  Returns the device tensor of
        per-iteration summed costs (no host sync)."""
        if self.step_count + n_iters > self.lr_t.numel():
            raise ValueError("max_steps exceeded")
        if self.mrf:
            if n_iters != 1:
                raise ValueError("spatial priors couple neighbouring voxels: one iteration per launch")
            return self._step_spatial()
        if n_iters > self.max_fuse:
            raise ValueError("at most %i fused iterations per launch" % self.max_fuse)
        e = self.engine_desc(row0=self.step_count % self.n_batches)
        ad = self.adam_desc(n_iters)
        cost_ptr = self.cost_hist.data_ptr() + 8 * self.step_count if want_cost else None
        L.check(self.lib.svbasl_step(C.byref(self.mdesc), C.byref(e), C.byref(ad), cost_ptr,
                                     self.nan_count.data_ptr(), _stream_ptr()))
        self.step_count += n_iters
        return self.cost_hist[self.step_count - n_iters:self.step_count]

    # ---- spatial prior: one launch per iteration ----
    def _prime_samples(self):
        """First iteration after the state was set from outside: the pre-pass kernel draws this iteration's samples
        of every local voxel, halo included (afterwards each step writes the next iteration's samples itself)."""
        if self.plan is not None and self.plan.world > 1:
            self.plan.exchange_halo(self.state)                # halo voxels' state from the adjacent ranks (set-up only)
        e = self.engine_desc()
        self.sample_spatial(e, self.step_count, self.sp_bufs[self.sp_cur])
        self.ak_grad.zero_()
        self.sp_valid = True
        if self.plan is not None and self.plan.world > 1:
            import torch.distributed as td
            torch.cuda.synchronize()
            td.barrier()                   # peer stores of the first iteration must not overtake a neighbour's pre-pass

    def _launch_spatial_iteration(self):
        """The launches of one spatial iteration for the current buffer parity (also what enable_graph captures).

        Default flow ("prepass"): step kernel -> pre-pass kernel for the NEXT iteration over the owned voxels (its
        samples of shard-boundary voxels also stored into the adjacent ranks' halo columns over NVLink) -> hyper step
        (all-reduce of the log-ak gradient over peer-memory mailboxes / NCCL, Adam on log ak, counter advance; the
        all-reduce is also the barrier that makes the mirrored samples visible).  The lightweight pre-pass kernel draws
        the samples at full occupancy for 17 us per million voxels; drawing them inside the step kernel ("fused",
        "separate_tail") costs 32 us and the per-CTA completion protocol of the fused tail another 11
        (profiles/r2_notes.md), so those flows are kept as measured alternatives only."""
        ad = self.adam_desc(1)
        mode = self.halo_mode if (self.plan is not None and self.plan.world > 1) else None
        flow = self.spatial_flow
        e = self.engine_desc(for_step=True)
        nxt = self.sp_bufs[1 - self.sp_cur]
        hy = None
        if flow == "prepass":
            es = self.engine_desc(for_step=True)               # descriptor of the sampler: keeps the peer pointers
            e.spatial_samples_out = None
            e.peer_lo = e.peer_hi = None
        elif flow == "fused" and mode in (None, "peer"):
            hy = self.hyper_desc()
        elif mode == "nccl":
            e.peer_lo = e.peer_hi = None
        L.check(self.lib.svbasl_step_spatial(C.byref(self.mdesc), C.byref(e), C.byref(ad),
                                             C.byref(hy) if hy is not None else None, self.cost_hist.data_ptr(),
                                             self.nan_count.data_ptr(), _stream_ptr()))
        if hy is not None:
            return
        if flow == "prepass":
            L.check(self.lib.svbasl_sample_spatial_next(C.byref(es), 1, nxt.data_ptr(), _stream_ptr()))
        if mode == "nccl":
            self.plan.exchange_halo(nxt.view(-1, self.ld))     # ncclSend/Recv of the boundary voxels' next samples
        if mode == "peer":                                     # mailbox all-reduce + barrier + log-ak step
            L.check(self.lib.svbasl_hyper_step_peers(
                self.log_ak.data_ptr(), self.ak_m.data_ptr(), self.ak_v.data_ptr(), self.ak_grad.data_ptr(),
                len(self.mrf), 1.0 / self.n_vox_global, self.lr_t.data_ptr(), self.step_dev.data_ptr(), self.b1, self.b2,
                self.adam_eps, self.plan.rank, self.plan.world, (C.c_void_p * self.plan.world)(*self._mail_ptrs),
                self.peer_status.data_ptr(), _stream_ptr()))
            return
        if mode is not None:
            self.reduce_fn(self.ak_grad)                       # NCCL all-reduce (also orders the peer stores)
        L.check(self.lib.svbasl_hyper_step_dev(self.log_ak.data_ptr(), self.ak_m.data_ptr(), self.ak_v.data_ptr(),
                                               self.ak_grad.data_ptr(), len(self.mrf), 1.0 / self.n_vox_global,
                                               self.lr_t.data_ptr(), self.step_dev.data_ptr(), self.b1, self.b2,
                                               self.adam_eps, _stream_ptr()))

    def _step_spatial(self):
        if not self.sp_valid:
            self._prime_samples()
        step = self.step_count
        if self.graphs is not None:
            self.graphs[self.sp_cur].replay()
        else:
            self._launch_spatial_iteration()
        self.sp_cur ^= 1
        self.step_count += 1
        return self.cost_hist[step:step + 1]

    def shard(self, plan, halo_mode="peer", reduce_fn=None):
        """Spatial prior sharded over the ranks of one box.  "peer" / "peer+nccl": both neighbour-sample buffers move
        into IPC-exportable device memory and the handles are exchanged with the adjacent ranks (once), so that the
        step kernel stores boundary voxels' samples directly into the neighbours' halo columns over NVLink; "peer"
        also exports a small mailbox per rank for the fused all-reduce of the log-ak gradient."""
        import torch.distributed as td
        if halo_mode not in ("peer", "peer+nccl", "nccl"):
            raise ValueError("halo_mode must be 'peer', 'peer+nccl' or 'nccl'")
        self.plan, self.halo_mode, self.reduce_fn = plan, halo_mode, reduce_fn
        if plan.world == 1 or halo_mode == "nccl":
            return
        if plan.world > L.MAX_PEERS:
            raise ValueError("peer-memory modes support at most %d ranks" % L.MAX_PEERS)
        n_bytes = 4 * self.sp_bufs[0].numel()
        handles = []
        for k in (0, 1):
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            with torch.cuda.device(self.dev):
                L.check(self.lib.svbasl_shared_alloc(n_bytes, C.byref(ptr), handle))
            view = _DevicePointer(ptr.value, tuple(self.sp_bufs[k].shape))
            t = torch.as_tensor(view, device=self.dev)
            t.copy_(self.sp_bufs[k])
            self.sp_bufs[k] = t
            self._shared.append((ptr, view))
            handles.append(bytes(handle))
        mine = {"rank": plan.rank, "ld": self.ld, "offset": plan.global_offset, "handles": handles}
        if halo_mode == "peer":
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            with torch.cuda.device(self.dev):
                L.check(self.lib.svbasl_shared_alloc(self.lib.svbasl_mailbox_bytes(plan.world), C.byref(ptr), handle))
            self._mailbox = ptr
            mine["mailbox"] = bytes(handle)
        everyone = [None] * plan.world
        td.all_gather_object(everyone, mine)
        self._opened = []

        def open_handle(raw):
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(raw)
            with torch.cuda.device(self.dev):                              # the ACCESSING device must be current
                L.check(self.lib.svbasl_shared_open(buf, C.byref(p)))
            self._opened.append(p.value)
            return p.value

        if halo_mode == "peer":
            self._mail_ptrs = [self._mailbox.value if r == plan.rank else open_handle(info["mailbox"])
                               for r, info in enumerate(everyone)]
        self.peers = {}
        for side, r in (("lo", plan.rank - 1), ("hi", plan.rank + 1)):
            if 0 <= r < plan.world:
                info = everyone[r]
                self.peers[side] = {"ptrs": [open_handle(h) for h in info["handles"]], "ld": info["ld"],
                                    "shift": plan.global_offset - info["offset"]}
        a = self.halo[0]
        lo_n, hi_n = plan.prev_halo_hi, plan.next_halo_lo      # owned voxels that lie in the neighbours' halos
        self._mirror = {"lo": (a, lo_n), "hi": (a + self.n_vox - hi_n, hi_n)}
        torch.cuda.synchronize()
        td.barrier()

    def enable_graph(self):
        """Capture the launch(es) of one spatial iteration as a CUDA graph per buffer parity; step() then costs one
        replay.  The iteration index lives in device memory (step_dev), so a replay needs no new arguments."""
        if not self.mrf:
            raise ValueError("graph replay is wired for the spatial-prior iteration (the others are one launch)")
        if not self.sp_valid:
            self._prime_samples()
        if self.reduce_fn is not None and self.halo_mode in ("peer+nccl", "nccl"):
            self.ak_grad.zero_()
            self.reduce_fn(self.ak_grad)                                  # NCCL communicator warm-up outside capture
        torch.cuda.synchronize()
        graphs = [None, None]
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        cur0 = self.sp_cur
        with torch.cuda.stream(side):
            for _k in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    self._launch_spatial_iteration()
                graphs[self.sp_cur] = g
                self.sp_cur ^= 1
        self.sp_cur = cur0
        torch.cuda.current_stream().wait_stream(side)
        self.graphs = graphs
        if self.plan is not None and self.plan.world > 1:
            import torch.distributed as td
            torch.cuda.synchronize()
            td.barrier()                  # the peer-memory reduction waits with a time-out: start the ranks together

    def check_peers(self):
        """Raise if a rank failed to arrive at the peer-memory all-reduce within its time-out (host sync)."""
        if self.mrf and int(self.peer_status.item()) != 0:
            raise L.SvbAslError("spatial iteration: a rank did not arrive at the all-reduce of the log-ak gradient "
                                "within the time-out; the results of this run are invalid")

    def release(self):
        """Tear down graph / peer-memory resources (before the process group is destroyed): captured graphs hold
        NCCL work, neighbours hold mappings of our shared buffers."""
        import torch.distributed as td
        torch.cuda.synchronize()
        self.graphs = None
        timed_out = bool(self.mrf) and int(self.peer_status.item()) != 0
        if self.peers is not None:
            for ptr in getattr(self, "_opened", []):
                self.lib.svbasl_shared_close(C.c_void_p(ptr))
            self._opened = []
            self.peers = None
            if td.is_available() and td.is_initialized():
                td.barrier()                       # every neighbour has closed its mapping of our buffers
            self.sp_bufs = [b.clone() for b in self.sp_bufs]
            for ptr, _view in self._shared:
                self.lib.svbasl_shared_free(ptr)
            self._shared = []
            if getattr(self, "_mailbox", None) is not None:
                self.lib.svbasl_shared_free(self._mailbox)
                self._mailbox = None
            self.halo_mode = "nccl"                # still usable, through NCCL
        torch.cuda.synchronize()
        if timed_out:
            raise L.SvbAslError("spatial iteration: a rank did not arrive at the all-reduce of the log-ak gradient "
                                "within the time-out; the results of this run are invalid")

    def finish(self):
        """Kept for callers of the round-1 API: every launch of an iteration is on the current stream."""

    def sample_spatial(self, e, step, out):
        """Pre-pass: theta samples of the spatially regularised parameters for all local voxels (halo included)."""
        L.check(self.lib.svbasl_sample_spatial(C.byref(e), self.ld, step, out.data_ptr(), _stream_ptr()))

    def elbo_grad(self, step=None, row0=0):
        """-> (cost [ld], grad [n_state, ld]) without updating anything."""
        e = self.engine_desc(row0=row0)
        cost = torch.zeros(self.ld, device=self.dev)
        grad = torch.zeros(self.n_state, self.ld, device=self.dev)
        step = self.step_count if step is None else step
        if self.mrf:
            # this evaluation's neighbour samples and log-ak gradient go to scratch buffers: the running ones belong
            # to the iteration in flight
            if self.plan is not None and self.plan.world > 1:
                self.plan.exchange_halo(self.state)
            scratch = torch.empty_like(self.sp_bufs[0])
            self.ak_grad_eval = torch.zeros_like(self.ak_grad)
            e.ak_grad = self.ak_grad_eval.data_ptr()
            e.spatial_samples = scratch.data_ptr()
            self.sample_spatial(e, step, scratch)
        L.check(self.lib.svbasl_elbo_grad(C.byref(self.mdesc), C.byref(e), step, cost.data_ptr(), grad.data_ptr(), None,
                                          _stream_ptr()))
        return cost, grad

    def model_fit(self):
        """Prediction at the posterior mean for every time point -> [T, ld]"""
        e = self.engine_desc()
        out = torch.zeros(self.T, self.ld, device=self.dev)
        L.check(self.lib.svbasl_model_fit(C.byref(self.mdesc), C.byref(e), out.data_ptr(), _stream_ptr()))
        return out

    def init_stats(self):
        """-> dict(mean_t, max_t, var_t, t_at_max), each [ld] (posterior initialisers, aslrest.py:461-520)"""
        tp = self.tpts
        if tp is None:
            tp = self.ti[:, None] + (self.zoff[None, :] if self.zoff is not None else 0.0)
            tp = tp.expand(self.T, self.ld).contiguous()
        return device_init_stats(self.data, tp)

    def fill_eps(self, step):
        eps = torch.zeros(self.N, self.S, self.ld, device=self.dev)
        L.check(self.lib.svbasl_fill_eps(eps.data_ptr(), self.ld, self.ld, self.vox_offset - self.halo[0], self.N,
                                         self.S, self.seed, step, _stream_ptr()))
        return eps

    # ---- results ----
    def posterior_mean(self):
        """Model-space posterior means [P, n_vox] (transform applied) and internal means/variances."""
        sl = slice(self.halo[0], self.halo[0] + self.n_vox)
        return self.state[:self.N, sl], torch.exp(self.state[self.N:2 * self.N, sl])


class HostFeeder:
    """
    Iterations fed from HOST memory, the way svb feeds every batch through ``feed_dict``: each ``step`` copies the
    batch's data rows (pinned host memory) to the device on a copy stream while the previous iteration computes,
    runs the fused step and copies the summed cost back (svbasl_step_host).  Time points travel either as a full
    [B, ld] array or in the model's low-rank form (the batch's TIs; the per-voxel slice offset stays resident).
    Spatial priors run through the same call (single GPU or peer-memory sharding: the iteration is one launch).
    """

    def __init__(self, fused):
        self.f = fused
        self.lib = fused.lib
        if fused.mrf and fused.plan is not None and fused.plan.world > 1 and fused.halo_mode != "peer":
            raise ValueError("host-fed spatial iterations over several ranks need halo_mode 'peer' (one launch)")
        self.ctx = C.c_void_p()
        L.check(self.lib.svbasl_host_ctx_create(C.byref(self.ctx), fused.ld, fused.B))
        self.cost = torch.zeros(2, dtype=torch.float64).pin_memory()
        self.calls = 0

    def step(self, host_data, host_tpts=None, host_ti=None, zoff_dev=None):
        """host_data [B, ld] (pinned); host_tpts [B, ld] (pinned) or host_ti [B] (pinned) + zoff_dev [ld] or None."""
        f = self.f
        hy = None
        if f.mrf:
            # host-fed spatial iterations use the single-launch form (next samples and hyper tail inside the step
            # kernel): svbasl_step_host is one launch per call
            if not f.sp_valid:
                f._prime_samples()
            torch.cuda.current_stream().synchronize() if self.calls == 0 else None
            hy = f.hyper_desc()
        e = f.engine_desc(for_step=bool(f.mrf))
        if host_tpts is None:
            e.tpts = None
            e.zoff = zoff_dev.data_ptr() if zoff_dev is not None else None
        ad = f.adam_desc(1)
        L.check(self.lib.svbasl_step_host(self.ctx, C.byref(f.mdesc), C.byref(e), C.byref(ad),
                                          C.byref(hy) if hy is not None else None, host_data.data_ptr(),
                                          host_tpts.data_ptr() if host_tpts is not None else None,
                                          host_ti.data_ptr() if host_ti is not None else None,
                                          self.cost.data_ptr() + 8 * (self.calls & 1)))
        if f.mrf:
            f.sp_cur ^= 1
        f.step_count += 1
        self.calls += 1

    def sync(self):
        L.check(self.lib.svbasl_host_sync(self.ctx))
        return float(self.cost[(self.calls - 1) & 1]) if self.calls else float("nan")

    def close(self):
        if self.ctx:
            self.lib.svbasl_host_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
