#!/usr/bin/env python
"""
bench.py - headline benchmark: voxel-iterations/second of the fused ELBO+gradient(+Adam) step, S samples.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], scripts/asl_example_sim.py + gen_test_data.py): aslrest multi-PLD pCASL,
6 PLDs, ftiss + delttiss + arterial component (fblood ARD, deltblood), S = 10 samples, sample-based latent
loss, 1,000,000 synthetic voxels PER GPU (weak scaling; voxels are independent so shards need no data-path
collective).  A "step" is one iteration over all voxels of the shard = one launch of the fused kernel.

One JSON line on rank 0; keys documented in DESIGN.md section 7.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PLDS = [0.25, 0.5, 0.75, 1.0, 1.25, 1.5]
# scripts/asl_example_sim.py:23-40 (+ arterial component, configs[1])
MODEL_OPTIONS = {"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "inferart": True}
FIT_OPTIONS = {"learning_rate": 0.05, "sample_size": 10, "force_num_latent_loss": True}
METRIC = "voxel-iters/sec (fused ELBO+grad, S samples)"


def synth_truth(n, seed):
    """gen_test_data.py:40-41 (ftiss~U(1,20), delttiss~U(0.6,2.5)) + an arterial component in 20% of voxels."""
    rng = np.random.default_rng(seed)
    ftiss = rng.uniform(1.0, 20.0, n)
    delt = rng.uniform(0.6, 2.5, n)
    fblood = rng.uniform(0.0, 10.0, n) * (rng.uniform(size=n) < 0.2)
    deltblood = np.maximum(delt - 0.3, 0.05)
    return np.stack([ftiss, delt, fblood, deltblood]).astype(np.float32), rng


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def cpu_port_rate(n_vox, n_iters, seed=20260101, threads=None):
    """The oracle port of the reference graph (oracle/svb_engine.py: op-for-op, one materialised [W,S,B]
    tensor per elementary op, torch autograd backward, TF-form Adam), float32, all host threads.
    TensorFlow itself is not installable in this image.  -> (voxel-iters/s, seconds per iteration)"""
    import torch
    from oracle import asl_models as om
    from oracle import svb_engine as eng
    from tests import helpers as H
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = om.AslConfig(casl=True, inferart=True, tau=1.8, t1b=1.65)
    spec = H.aslrest_spec(cfg)
    truth, rng = synth_truth(n_vox, seed)
    tis = np.asarray([1.8 + p for p in PLDS])
    t = torch.as_tensor(np.repeat(tis[:, None], n_vox, 1), dtype=torch.float32)
    par = [torch.as_tensor(truth[i]).reshape(n_vox, 1, 1) for i in range(4)]
    data = om.evaluate(cfg, par, t.T.unsqueeze(1))[:, 0, :].T + torch.randn(6, n_vox)
    dnp = data.numpy()
    state = eng.initial_state(spec, [np.maximum(dnp.mean(0), 0.1), 1.3, np.maximum(dnp.max(0), 0.1), 1.3,
                                     np.log(np.maximum(dnp.var(0), 1.0))], [1.5, 1.0, 1.5, 1.0, 1.02], n_vox,
                              dtype=torch.float32)
    hyper = torch.zeros(0)
    opt = eng.Adam(lr=0.05)
    times = []
    warm = 2                                  # thread pool / allocator warm-up iterations, not timed
    for it in range(n_iters + warm):
        t0 = time.perf_counter()
        eps = torch.randn(5, 10, n_vox)
        cost, gs, _gh, _ = eng.cost_and_grad(spec, state, hyper, data, t, eps)
        opt.update({"state": (state, gs)})
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times[warm:]))
    return n_vox / sec, sec, threads


def run_reference(args, out):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_vox = args.cpu_voxels
    rate, sec, threads = cpu_port_rate(n_vox, max(1, min(args.steps, 20)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "voxel-iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "asl_example_sim: aslrest multi-PLD pCASL, ftiss+delttiss+arterial, S=10, B=T=6",
                   "voxels_per_step": n_vox, "note": "bounded sample of the 1M-voxel workload"},
        "cpu_baseline": {"value": rate, "unit": "voxel-iters/s", "cores": threads, "kind": "port",
                         "sample": "%d voxels x %d iterations; PyTorch-CPU op-for-op restatement of the reference "
                                   "TF graph (TensorFlow/svb not installable in this image)" % (n_vox, args.steps)},
        "e2e": {"value": rate, "unit": "voxel-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()


# ----------------------------------------------------------------------------------------------------
def _claim_stdout():
    """Library chatter (e.g. NCCL's version banner) goes to stderr; stdout carries exactly ONE JSON line."""
    real = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--voxels", type=int, default=1_000_000, help="voxels per GPU")
    ap.add_argument("--cpu-voxels", type=int, default=100_000, help="voxels of the bounded CPU sample")
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import torch.distributed as td
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200 import _lib as L
    from svb_models_asl_b200.svbcompat.fit import SvbFit

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    W, K, WU = args.voxels, args.steps, max(3, args.warmup)

    # ---- synthetic shard (gen_test_data.py restated; generated through the plugin's own evaluate kernel) ----
    truth, rng = synth_truth(W, 20260101 + rank)
    dm0 = DataModel(np.zeros((1, len(PLDS)), dtype=np.float32))
    gen = AslRestModel(dm0, **{**MODEL_OPTIONS, "t1b": 1.6})            # generator t1b=1.6 (gen_test_data.py:28)
    tis = np.asarray(gen.tis, dtype=np.float32)
    sig = gen.evaluate(list(truth.reshape(4, W, 1, 1)), tis.reshape(1, 1, -1))[:, 0, :]
    sig = sig + torch.randn(sig.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + rank))
    data_host = sig.cpu().numpy()                                       # [W, T]
    dm = DataModel(data_host)
    model = AslRestModel(dm, **MODEL_OPTIONS)                           # fit model: default t1b=1.65
    fit = SvbFit(dm, model, **FIT_OPTIONS)
    fit.lo, fit.hi = 0, W                                               # every rank owns its own W voxels (weak scaling)
    fit._setup(model.tpts(), dm.data_flattened, None, FIT_OPTIONS["sample_size"], FIT_OPTIONS["learning_rate"],
               epochs=4 * (K + WU) + 64, force_num_latent_loss=FIT_OPTIONS["force_num_latent_loss"])
    f = fit.fused
    f.n_vox_global = W * world
    n_state = f.n_state
    bytes_per_voxel = 8 * f.B + 24 * n_state          # data+tpts read; state/m/v read + written
    lane_instr_per_voxel = 60 * 50 + 10 * 60 + 300    # FP32 lane-instructions per voxel-iteration

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K launches, CUDA events on the launch stream ----
    for _ in range(WU):
        f.step(1)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record()
    for i in range(K):
        f.step(1)
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[K])
    per_launch_ms = float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(K)]))
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = W * world * K / (total_ms * 1e-3)
    final_cost = float(f.cost_hist[f.step_count - 1].item()) / W
    assert math.isfinite(final_cost), "non-finite cost"

    # ---- end to end through the C ABI with HOST buffers (svb feeds each batch via feed_dict) ----
    lib = L.load()
    ctx = C.c_void_p()
    L.check(lib.svbasl_host_ctx_create(C.byref(ctx), f.ld, f.B))
    h_data = torch.from_numpy(np.ascontiguousarray(data_host.T)).pin_memory()          # [B, ld]
    h_tpts = torch.from_numpy(np.ascontiguousarray(model.tpts().T)).pin_memory()
    h_cost = torch.zeros(2, dtype=torch.float64).pin_memory()

    def host_step(i):
        e = f.engine_desc()
        ad = f.adam_desc(1)
        L.check(lib.svbasl_step_host(ctx, C.byref(f.mdesc), C.byref(e), C.byref(ad), h_data.data_ptr(),
                                     h_tpts.data_ptr(), h_cost.data_ptr() + 8 * (i & 1)))
        f.step_count += 1

    for i in range(WU):
        host_step(i)
    L.check(lib.svbasl_host_sync(ctx))
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_step(i)
    L.check(lib.svbasl_host_sync(ctx))
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(te, op=td.ReduceOp.MAX)
    e2e_s = float(te.item())
    assert math.isfinite(float(h_cost[0])) and math.isfinite(float(h_cost[1]))
    L.check(lib.svbasl_host_ctx_destroy(ctx))
    h2d = 2 * 4 * f.B * f.ld

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    peak, peak_src, _ = peaks()
    achieved = bytes_per_voxel * W / (per_launch_ms * 1e-3) / 1e9
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    fp32_peak = 148 * 128 * sm_hz
    line = {
        "metric": METRIC, "value": value, "unit": "voxel-iters/s", "n_gpus": world, "steps": K, "warmup": WU,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "asl_example_sim: aslrest multi-PLD pCASL (6 PLD), ftiss+delttiss+arterial, S=10, "
                               "B=T=6, sample-based latent loss, Adam fused",
                   "voxels_per_gpu": W, "n_state": n_state, "rng": "philox4x32-10 in-kernel",
                   "l2": "working set %.0f MB per step > 126 MB L2 (no flush needed)" % (bytes_per_voxel * W / 1e6)},
        "clocks": clocks,
        "e2e": {"value": W * world * K / e2e_s, "unit": "voxel-iters/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_s / K * 1e3,
                "path": "svbasl_step_host: pinned host batch -> H2D -> fused step -> D2H cost, double-buffered"},
        "gpu_launches": K,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "step_kernel<AslRest<0x7>,6,0>",
                     "algorithmic_bytes_per_voxel_iter": bytes_per_voxel, "avg_launch_ms": per_launch_ms,
                     "fp32": {"lane_instr_per_voxel_iter": lane_instr_per_voxel,
                              "achieved_tlane_per_s": lane_instr_per_voxel * W / (per_launch_ms * 1e-3) / 1e12,
                              "peak_tlane_per_s": fp32_peak / 1e12,
                              "frac": lane_instr_per_voxel * W / (per_launch_ms * 1e-3) / fp32_peak,
                              "note": "FP32-pipe roofline 148 SM x 128 lanes x measured SM clock (SURVEY 8d)"}},
        "final_mean_cost": final_cost,
    }
    if not args.no_cpu_baseline:
        rate, sec, threads = cpu_port_rate(args.cpu_voxels, args.cpu_iters)
        line["cpu_baseline"] = {"value": rate, "unit": "voxel-iters/s", "cores": threads, "kind": "port",
                                "sample": "%d voxels x %d iterations (%.1f s/iter); PyTorch-CPU op-for-op port of the "
                                          "reference TF graph" % (args.cpu_voxels, args.cpu_iters, sec)}
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
