"""
``SvbFit``: the inference engine the reference drives through ``svb.main.run`` - posterior, noise model,
priors, cost and Adam training loop - re-built around the fused
B200 kernel: one launch per iteration does what svb's ``sess.run(optimize)`` does for the ASL plugins.

The svb source is not part of the reference tree; the conventions followed here are listed in DESIGN.md
section 5 with a switch for each (``cov_convention``, ``ard_phi_max``, ``init_delt_var`` ...).

Voxels are sharded contiguously over the ranks of ``torch.distributed`` when it is initialised (one process
per GPU); voxel-wise priors need no data-path collective, only the scalar cost is all-reduced for reporting.
"""
import math
import time

import numpy as np
import torch

from . import dist as _dist
from .parameter import Parameter
from .utils import LogBase
from ..ops import FusedSvb, InitData, device_array
from ..sharding import ShardPlan, shard_bounds, world_info


def noise_parameter():
    """svb's NoiseParameter: prior LogNormal(1, 2e5), posterior LogNormal(1, 1.02), initialised from the
    per-voxel data variance floored at 1."""
    def _init(_param, _t, data):
        st = getattr(data, "device_stats", None)
        var = np.asarray(st["var_t"]) if st is not None else np.asarray(data).var(axis=1)
        return np.where(var < 1, 1.0, var).astype(np.float32), None
    return Parameter("noise", prior=_dist.LogNormal(1.0, 2e5), post=_dist.LogNormal(1.0, 1.02), post_init=_init)


_dist_world = world_info


class SvbFit(LogBase):
    def __init__(self, data_model, fwd_model, **kwargs):
        LogBase.__init__(self)
        self.data_model = data_model
        self.model = fwd_model
        self.params = list(fwd_model.params) + [noise_parameter()]
        self.n_params = len(self.params)
        self.rank, self.world = _dist_world()
        self.lo, self.hi = shard_bounds(data_model.n_nodes, self.rank, self.world)
        self.opts = kwargs
        self.fused = None

    # ------------------------------------------------------------------
    def _setup(self, tpts, data, batch_size, sample_size, learning_rate, epochs, **kwargs):
        lo, hi = self.lo, self.hi
        data = np.asarray(data, dtype=np.float32)
        tpts = np.asarray(tpts, dtype=np.float32)
        if tpts.ndim == 1:
            tpts = tpts[None, :]
        if tpts.shape[0] == 1:
            tpts = np.broadcast_to(tpts, data.shape)
        n_t = data.shape[1]
        force_num = bool(kwargs.get("force_num_latent_loss", False))
        all_normal = all(isinstance(p.prior_dist, _dist.Normal) and p.prior_type in ("N", "A") for p in self.params)
        latent = "analytic" if (all_normal and not force_num) else "numeric"
        prior_types = [p.prior_type for p in self.params]
        neighbours = None
        halo = (0, 0)
        self.plan = ShardPlan(self.data_model.n_nodes, self.rank, self.world,
                              self.data_model.neighbour_table if "M" in prior_types else None)
        if (self.lo, self.hi) != (self.plan.lo, self.plan.hi):           # caller overrode the shard (bench.py)
            self.plan = ShardPlan(hi - lo, 0, 1, None)
            self.plan.lo, self.plan.hi = lo, hi
        plan = self.plan
        if "M" in prior_types:
            neighbours = plan.neighbours_local                           # [6, ld], local indices
            halo = (plan.halo_lo, plan.halo_hi)
        g0, g1 = lo - halo[0], hi + halo[1]                              # local arrays cover [g0, g1) of the global order
        n_batches = int(math.ceil(n_t / (batch_size or n_t)))
        self.fused = FusedSvb(
            self.model, data[g0:g1].T.copy(), tpts[g0:g1].T.copy(), n_samples=sample_size,
            batch_size=batch_size or n_t, latent=latent, cov_llt=(kwargs.get("cov_convention", "LtL") == "LLt"),
            learning_rate=learning_rate, seed=int(kwargs.get("seed", 1)), prior_types=prior_types,
            prior_means=[np.mean(p.prior_dist.mean) for p in self.params],
            prior_vars=[np.mean(p.prior_dist.var) for p in self.params],
            n_vox_global=self.data_model.n_nodes, vox_offset=lo, halo=halo, neighbours=neighbours,
            ak_init=float(kwargs.get("ak", 1e-5)), ard_phi_max=kwargs.get("ard_phi_max", 1e6),
            latent_weight=float(kwargs.get("latent_weight", 1.0)), max_steps=epochs * n_batches + 1)
        own = slice(halo[0], halo[0] + (hi - lo))
        stats = None
        if kwargs.get("device_init", True):
            stats = {k: v[own].cpu().numpy() for k, v in self.fused.init_stats().items()}
        local_data = InitData(data[lo:hi], stats)
        local_data.voxel_slice = slice(lo, hi)
        means, variances = self._initial_posterior(tpts[lo:hi], local_data)
        self.fused.set_posterior(means, variances)
        if "M" in prior_types:
            # "peer": next-iteration samples of boundary voxels stored straight into the neighbours' halos over NVLink
            # peer memory, log-ak all-reduce over peer-memory mailboxes (the iteration is replayed as a CUDA graph);
            # "peer+nccl": the all-reduce by NCCL; "nccl": send/recv of the halo samples
            mode = kwargs.get("halo_mode", "peer")
            if self.world > 1:
                self.fused.shard(plan, halo_mode=mode, reduce_fn=ShardPlan.allreduce_sum)
            if kwargs.get("use_graph", True) and (self.world == 1 or mode in ("peer", "peer+nccl")):
                self.fused.enable_graph()

    def _initial_posterior(self, local_t, local_data):
        """svb: post_init(param, t, data) -> (mean, var) in MODEL space, mapped to internal values.  Each rank
        initialises its own shard; the per-voxel reductions of the data (mean / max / variance / time of the maximum
        over time, aslrest.py:467,490,497-501) come from the device (svbasl_init_stats) and travel with the `data`
        argument (ops.InitData), so nothing of size [W, T] is reduced on the host."""
        n = len(local_data)
        means, variances = [], []
        for p in self.params:
            mean, var = None, None
            if p.post_init is not None:
                mean, var = p.post_init(p, local_t, local_data)
                if mean is not None:
                    mean = p.post_dist.transform.int_values(np.asarray(mean, dtype=np.float32))
            if mean is None:
                mean = np.full(n, np.mean(p.post_dist.mean), dtype=np.float32)
            if var is None:
                var = np.full(n, np.mean(p.post_dist.var), dtype=np.float32)
            means.append(np.broadcast_to(np.asarray(mean, dtype=np.float32), (n,)))
            variances.append(np.broadcast_to(np.asarray(var, dtype=np.float32), (n,)))
        return means, variances

    def setup_from_device(self, data_dev, plan, ti, zoff_dev, batch_size, sample_size, learning_rate, max_steps,
                          **kwargs):
        """The same set-up as train() performs, for data that already live on this rank's GPU: `data_dev` [T, ld]
        covers the rank's local voxel range of `plan` (halo included), time points in the model's low-rank form
        (`ti` [T] + `zoff_dev` [ld], aslrest.py:438-440).  Used for volumes too large to stage through host arrays
        (bench.py's 10 M-voxel spatial fit); everything downstream (FusedSvb, posterior initialisers, sharding) is
        the code path of _setup()."""
        self.plan = plan
        self.lo, self.hi = plan.lo, plan.hi
        prior_types = [p.prior_type for p in self.params]
        spatial = "M" in prior_types
        halo = (plan.halo_lo, plan.halo_hi) if spatial else (0, 0)
        self.fused = FusedSvb(
            self.model, data_dev, None, ti=ti, zoff=zoff_dev, n_samples=sample_size, batch_size=batch_size,
            latent="numeric", cov_llt=(kwargs.get("cov_convention", "LtL") == "LLt"), learning_rate=learning_rate,
            seed=int(kwargs.get("seed", 1)), prior_types=prior_types,
            prior_means=[np.mean(p.prior_dist.mean) for p in self.params],
            prior_vars=[np.mean(p.prior_dist.var) for p in self.params], n_vox_global=plan.n_global, vox_offset=plan.lo,
            halo=halo, neighbours=plan.neighbours_local if spatial else None, ak_init=float(kwargs.get("ak", 1e-5)),
            ard_phi_max=kwargs.get("ard_phi_max", 1e6), latent_weight=float(kwargs.get("latent_weight", 1.0)),
            max_steps=max_steps)
        own = slice(halo[0], halo[0] + plan.n_own)
        stats = {k: v[own].cpu().numpy() for k, v in self.fused.init_stats().items()}
        local = InitData(np.empty((plan.n_own, 0), dtype=np.float32), stats)
        local.voxel_slice = slice(plan.lo, plan.hi)
        self.fused.set_posterior(*self._initial_posterior(None, local))
        if spatial:
            mode = kwargs.get("halo_mode", "peer")
            if plan.world > 1:
                self.fused.shard(plan, halo_mode=mode, reduce_fn=ShardPlan.allreduce_sum)
            if kwargs.get("use_graph", True) and (plan.world == 1 or mode in ("peer", "peer+nccl")):
                self.fused.enable_graph()
        return self.fused

    # ------------------------------------------------------------------
    def train(self, tpts, data, batch_size=None, epochs=100, learning_rate=0.1, sample_size=None, display_step=1,
              iters_per_launch=None, **kwargs):
        """Returns the training history dict (mean_cost per epoch; voxel cost / parameter histories when the
        corresponding save_* option is set).

        iters_per_launch: iterations fused into one kernel launch (voxel-wise priors; svbasl_adam.n_iters).  Default:
        16, lowered to the number of batches per epoch when a per-epoch history is recorded (save_cost_history /
        save_param_history need the state at every epoch boundary)."""
        sample_size = sample_size or 5
        self._setup(tpts, data, batch_size, sample_size, learning_rate, epochs, **kwargs)
        f = self.fused
        n_batches = f.n_batches
        want_vc = bool(kwargs.get("save_cost_history", False))
        want_ph = bool(kwargs.get("save_param_history", False))
        hist = {"mean_cost": np.zeros(epochs + 1, dtype=np.float64)}
        cost_dev = torch.zeros(epochs + 1, device=f.dev, dtype=torch.float64)
        vc = torch.zeros(epochs + 1, f.n_vox, device=f.dev) if want_vc else None
        ph = torch.zeros(epochs + 1, f.N, f.n_vox, device=f.dev) if want_ph else None
        sl = slice(f.halo[0], f.halo[0] + f.n_vox)
        stream = kwargs.get("log_stream") if self.rank == 0 else None

        def record(epoch):
            if want_vc or epoch == 0:
                c = self.full_cost()
                cost_dev[epoch] = c.sum(dtype=torch.float64)
                if want_vc:
                    vc[epoch] = c
            if want_ph:
                ph[epoch] = f.state[:f.N, sl]

        def log(epoch):
            if stream and display_step and epoch % max(1, display_step) == 0:
                # one host sync per displayed epoch; keep display_step large for big fits
                mc = float(f.cost_hist[(epoch - 1) * n_batches:epoch * n_batches].mean()) / f.n_vox
                stream.write(" - Epoch %04d: mean cost=%f (shard of %i voxels)\n" % (epoch, mc, f.n_vox))

        record(0)
        t0 = time.time()
        total = epochs * n_batches
        per_epoch = want_vc or want_ph
        fuse = 1 if f.mrf else max(1, min(int(iters_per_launch or 16), f.max_fuse))
        done = 0
        while done < total:
            # a launch never crosses an epoch boundary at which something is recorded or displayed
            if per_epoch:
                stop = (done // n_batches + 1) * n_batches
            elif stream and display_step:
                span = n_batches * max(1, display_step)
                stop = min(total, (done // span + 1) * span)
            else:
                stop = total
            k = min(fuse, stop - done)
            f.step(k)
            done += k
            if done % n_batches == 0:
                epoch = done // n_batches
                if per_epoch:
                    record(epoch)
                log(epoch)
                if f.mrf and self.world > 1 and epoch % 64 == 0:
                    f.check_peers()                             # a lost rank surfaces here, not at teardown
        if f.mrf:
            f.check_peers()
        if not want_vc:
            # mean of each epoch's batch costs, from the per-iteration sums the kernels left on the device
            cost_dev[1:] = f.cost_hist[:total].view(epochs, n_batches).mean(dim=1)
        torch.cuda.synchronize()
        self.runtime = time.time() - t0
        total_cost = cost_dev.clone()
        if self.world > 1:
            import torch.distributed as td
            td.all_reduce(total_cost)                           # the only collective: scalar cost, for reporting
        hist["mean_cost"] = (total_cost / self.data_model.n_nodes).cpu().numpy()
        if want_vc:
            hist["voxel_cost"] = vc.T.cpu().numpy()
        if want_ph:
            hist["params"] = ph.permute(2, 0, 1).cpu().numpy()  # [W, epochs+1, P']
        hist["nan_skips"] = int(f.nan_count.item())
        return hist

    # ------------------------------------------------------------------
    def full_cost(self):
        """Per-voxel cost on ALL time points (what svb reports per epoch) -> [n_vox] device tensor"""
        f = self.fused
        saved = (f.B, f.n_batches)
        f.B, f.n_batches = f.T, 1
        try:
            cost, _ = f.elbo_grad(row0=0)
        finally:
            f.B, f.n_batches = saved
        return cost[f.halo[0]:f.halo[0] + f.n_vox]

    def model_moments(self):
        """Posterior mean / variance of every parameter in MODEL space (svb's model_means / model_vars:
        exp(.) of both moments for LogNormal, |mean| for FoldedNormal) -> numpy [P', n_vox] each"""
        mean, var = self.fused.posterior_mean()
        mean, var = mean.clone(), var.clone()
        for i, p in enumerate(self.params):
            code = p.post_dist.transform.code
            if code == _dist.XF_EXP:
                mean[i], var[i] = torch.exp(mean[i]), torch.exp(var[i])
            elif code == _dist.XF_ABS:
                mean[i] = torch.abs(mean[i])
        return mean.cpu().numpy(), var.cpu().numpy()

    def model_fit(self):
        f = self.fused
        return f.model_fit()[:, f.halo[0]:f.halo[0] + f.n_vox].T.cpu().numpy()

    def close(self):
        """Release graph / peer-memory resources of a sharded fit (results stay readable)."""
        if self.fused is not None:
            self.fused.release()

    def gather(self, local):
        """Concatenate per-shard [.., n_vox] (last axis = voxels... first axis here) arrays on every rank."""
        if self.world == 1:
            return local
        import torch.distributed as td
        parts = [None] * self.world
        td.all_gather_object(parts, local)
        return np.concatenate(parts, axis=0)
