"""
The N > 1 host path on CPU: two gloo ranks shard a masked volume, exchange halo state and all-reduce the global
sums; the result must equal the single-process computation.  The per-voxel arithmetic is stood in for by the
host build of the device code (tests/hostsim) so the same sharded sequence the GPUs run - pre-pass over
owned+halo voxels, main step over owned voxels, ak-gradient all-reduce - is exercised end to end without a GPU.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    from oracle import asl_models as om
    from tests import helpers as H
    from svb_models_asl_b200.svbcompat.data import DataModel
    rng = np.random.default_rng(77)
    shape = (6, 5, 4)
    mask = (rng.uniform(size=shape) < 0.8).astype(np.int16)
    dm = DataModel(np.zeros(shape + (6,), dtype=np.float32), mask=mask)
    W = dm.n_nodes
    cfg = om.AslConfig(casl=True, tau=1.8, t1b=1.65)
    spec = H.aslrest_spec(cfg, mrf=(0,))
    prob = H.synth_problem(cfg, spec, W, rng)
    return cfg, spec, prob, dm.neighbour_table(), W


def _run_local(plan, cfg, spec, prob, log_ak, seed, step):
    """One ELBO+gradient evaluation of this rank's shard through the host build; returns owned cost, grad, ak sum."""
    from tests import helpers as H
    be = H.Backend("hostsim")
    state = plan.take(prob["state"], axis=1)
    data = plan.take(prob["data"], axis=1)
    tpts = plan.take(prob["tpts"], axis=1)
    m = be.model_desc(cfg)
    e, bufs = be.engine_desc(spec, state, data, tpts, None, seed=seed, neighbours=plan.neighbours_local,
                             log_ak=log_ak, w_begin=plan.halo_lo, n_vox=plan.n_own, vox_offset=plan.global_offset,
                             n_vox_global=plan.n_global)
    return be, m, e, bufs


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as td
    td.init_process_group("gloo", rank=rank, world_size=world)
    from svb_models_asl_b200.sharding import ShardPlan
    cfg, spec, prob, nb, W = _problem()
    plan = ShardPlan(W, rank, world, nb)
    log_ak = np.asarray([-1.0], dtype=np.float32)
    # scramble the halo columns, then restore them through the exchange
    state = torch.as_tensor(plan.take(prob["state"], axis=1).copy())
    truth = state.clone()
    state[:, :plan.halo_lo] = -999.0
    state[:, plan.halo_lo + plan.n_own:] = -999.0
    plan.exchange_halo(state)
    assert torch.equal(state, truth), "halo exchange did not reproduce the neighbours' state"
    be, m, e, bufs = _run_local(plan, cfg, spec, prob, log_ak, seed=3, step=5)
    be.sample_spatial(e, bufs, step=5)
    cost, grad, _ = be.elbo_grad(m, e, spec.n_state, step=5, nbt=6)
    ak = torch.as_tensor(be.get(bufs["ak_grad"])[:1].copy())
    csum = torch.tensor([cost[plan.own].astype(np.float64).sum()])
    ShardPlan.allreduce_sum(ak)
    ShardPlan.allreduce_sum(csum)
    gathered_cost = plan.gather_owned(cost[plan.own])
    gathered_grad = plan.gather_owned(grad[:, plan.own], axis=1)
    if rank == 0:
        np.savez(out, cost=gathered_cost, grad=gathered_grad, ak=ak.numpy(), csum=csum.numpy(),
                 halos=np.asarray([plan.halo_lo, plan.halo_hi, plan.next_halo_lo]))
    td.destroy_process_group()


def test_two_rank_spatial_step_equals_single_process(tmp_path):
    from svb_models_asl_b200.sharding import ShardPlan
    out = str(tmp_path / "sharded.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    cfg, spec, prob, nb, W = _problem()
    plan = ShardPlan(W, 0, 1, nb)
    assert plan.halo_lo == 0 and plan.halo_hi == 0 and plan.ld == W
    be, m, e, bufs = _run_local(plan, cfg, spec, prob, np.asarray([-1.0], dtype=np.float32), seed=3, step=5)
    be.sample_spatial(e, bufs, step=5)
    cost, grad, _ = be.elbo_grad(m, e, spec.n_state, step=5, nbt=6)
    ak = be.get(bufs["ak_grad"])[:1]
    # identical Philox draws per GLOBAL voxel -> the sharded run reproduces the single-process one
    np.testing.assert_allclose(got["cost"], cost, rtol=1e-6)
    np.testing.assert_allclose(got["grad"], grad, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(got["ak"], ak, rtol=1e-6)
    assert got["csum"][0] == pytest.approx(cost.astype(np.float64).sum(), rel=1e-9)
    assert got["halos"][1] > 0 and got["halos"][2] > 0        # rank 0 has an upper halo, rank 1 a lower one


def test_shard_bounds_cover_everything():
    from svb_models_asl_b200.sharding import shard_bounds
    for n in (0, 1, 7, 33222, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_rejects_too_thin_shards():
    from svb_models_asl_b200.sharding import ShardPlan
    from svb_models_asl_b200.svbcompat.data import DataModel
    dm = DataModel(np.zeros((4, 6, 6, 1), dtype=np.float32))
    nb = dm.neighbour_table()
    with pytest.raises(ValueError):
        ShardPlan(dm.n_nodes, 1, 8, nb)                        # 18 voxels per rank < one 36-voxel plane
