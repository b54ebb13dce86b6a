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

# The headline workload is BASELINE.json configs[1] ("sim_art").  The others are the remaining configs, kept
# for our own measurements (`--workload`); the driver only runs the default.
WORKLOADS = {
    "sim_art": dict(model="aslrest", options=MODEL_OPTIONS, batch=None, repeats=1, voxels=1_000_000,
                    desc="asl_example_sim: aslrest multi-PLD pCASL (6 PLD), ftiss+delttiss+arterial, S=10, B=T=6"),
    "real_like": dict(model="aslrest", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8], "slicedt": 0.0452},
                      batch=6, repeats=8, voxels=1_000_000,
                      desc="asl_example: aslrest 6 PLD x 8 repeats, slicedt, ftiss+delttiss, S=10, T=48, B=6"),
    "disp": dict(model="aslrest_disp", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8], "inferart": True},
                 batch=6, repeats=8, voxels=250_000,
                 desc="aslrest_disp gamma dispersion, 6 PLD x 8 repeats, ftiss+delttiss+arterial+s+sp, S=10, T=48, B=6"),
    "nn": dict(model="aslnn", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1]}, batch=None, repeats=1,
               voxels=1_000_000, desc="aslnn MLP surrogate 2-10-10-1, ftiss+delttiss, S=10, B=T=6"),
    "spatial": dict(model="aslrest", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8],
                                              "param_overrides": {"ftiss": {"prior_type": "M"}}},
                    batch=6, repeats=8, voxels=1_000_000, cube=True,
                    desc="aslrest with spatial MRF prior on ftiss, 6 PLD x 8 repeats, S=10, T=48, B=6, 100^3 volume"),
}


def synth_truth(n, seed):
    """gen_test_data.py:40-41 (ftiss~U(1,20), delttiss~U(0.6,2.5)) + an arterial component in 20% of voxels."""
    rng = np.random.default_rng(seed)
    ftiss = rng.uniform(1.0, 20.0, n)
    delt = rng.uniform(0.6, 2.5, n)
    fblood = rng.uniform(0.0, 10.0, n) * (rng.uniform(size=n) < 0.2)
    deltblood = np.maximum(delt - 0.3, 0.05)
    return np.stack([ftiss, delt, fblood, deltblood]).astype(np.float32), rng


def measured_traffic(workload, n_vox):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per voxel of that capture), scaled to
    this launch's voxel count; None when no capture exists for the workload."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    entry = json.load(open(path)).get(workload)
    return None if not entry else entry["dram_bytes_per_voxel"] * n_vox


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
def prefer_gpu_local_host_memory(dev_index):
    """Best effort, Linux only: ask the kernel to place this process's NEW host pages (the pinned staging buffers
    of the end-to-end leg) on the NUMA node the GPU hangs off, and run on that node's CPUs when the cpuset allows.
    Returns a short description for the JSON line.  Never fatal.  Undone by restore_host_placement() once the
    buffers exist (the CPU baseline must see every core)."""
    global _SAVED_AFFINITY
    _SAVED_AFFINITY = os.sched_getaffinity(0)
    try:
        import torch
        props = torch.cuda.get_device_properties(dev_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        if node < 0:
            return "gpu numa node unknown"
        note = "gpu %s on numa node %d" % (bdf, node)
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
            note += ", cpus pinned to %d local" % len(cpus & allowed)
        else:
            note += ", no local cpu in cpuset"
        mask = C.c_ulong(1 << node)
        libc = C.CDLL(None, use_errno=True)
        rc = libc.syscall(238, 1, C.byref(mask), C.c_ulong(8 * C.sizeof(C.c_ulong)))    # set_mempolicy(MPOL_PREFERRED)
        note += ", mempolicy preferred" if rc == 0 else ", mempolicy refused (errno %d)" % C.get_errno()
        return note
    except Exception as exc:                                                       # noqa: BLE001
        return "numa placement skipped: %s" % (str(exc)[:80],)


_SAVED_AFFINITY = None


def restore_host_placement():
    try:
        if _SAVED_AFFINITY:
            os.sched_setaffinity(0, _SAVED_AFFINITY)
        C.CDLL(None).syscall(238, 0, None, C.c_ulong(0))                           # MPOL_DEFAULT
    except Exception:                                                              # noqa: BLE001
        pass


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
    ap.add_argument("--voxels", type=int, default=None, help="voxels per GPU")
    ap.add_argument("--workload", default="sim_art", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-voxels", type=int, default=100_000, help="voxels of the bounded CPU sample")
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--iters-per-launch", type=int, default=8,
                    help="additional measurement: this many iterations fused per launch (state, data and Adam "
                         "moments stay on chip between them); 0 = skip")
    ap.add_argument("--halo-mode", default="peer", choices=["peer", "peer+nccl", "nccl"],
                    help="spatial workload on N > 1 GPUs: how halo state and the log-ak gradient travel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import torch.distributed as td
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200 import _lib as L
    from svb_models_asl_b200.plugin import get_model_class
    from svb_models_asl_b200.svbcompat.fit import SvbFit

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    W, K, WU = args.voxels or wl["voxels"], args.steps, max(3, args.warmup)
    if wl.get("cube"):
        side = int(round(W ** (1.0 / 3.0)))
        W = side ** 3

    # ---- synthetic shard (gen_test_data.py restated; generated through the plugin's own evaluate kernel) ----
    cube = bool(wl.get("cube"))            # one shared volume, sharded over the ranks (strong scaling)
    truth, rng = synth_truth(W, 20260101 + (0 if cube else rank))
    reps = wl["repeats"]
    dm0 = DataModel(np.zeros((1, len(PLDS) * reps), dtype=np.float32))
    # generator: aslrest tissue + arterial with t1b=1.6 (gen_test_data.py:28), whatever model is then fitted
    gen = AslRestModel(dm0, **{**MODEL_OPTIONS, "repeats": [reps], "t1b": 1.6})
    tis = np.repeat(np.asarray(gen.tis, dtype=np.float32), reps)
    sig = gen.evaluate(list(truth.reshape(4, W, 1, 1)), tis.reshape(1, 1, -1))[:, 0, :]
    sig = sig + torch.randn(sig.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + (0 if cube else rank)))
    data_host = sig.cpu().numpy()                                       # [W, T]
    del sig
    if wl.get("cube"):
        data_host = data_host.reshape(side, side, side, -1)
    dm = DataModel(data_host)
    model_opts = dict(wl["options"])
    if wl["model"] == "aslnn":
        wdir = os.path.join(ROOT, "trained_data")
        if not os.path.exists(os.path.join(wdir, "weights0.npy")):
            raise SystemExit("aslnn workload needs trained_data/ (python scripts/retrain_model.py)")
        model_opts["train_load"] = wdir
    model = get_model_class(wl["model"])(dm, **model_opts)              # fit model: default t1b=1.65
    fit = SvbFit(dm, model, **FIT_OPTIONS)
    if not wl.get("cube"):
        fit.lo, fit.hi = 0, W                                           # every rank owns its own W voxels (weak scaling)
    fit._setup(model.tpts(), dm.data_flattened, wl["batch"], FIT_OPTIONS["sample_size"], FIT_OPTIONS["learning_rate"],
               epochs=4 * (K + WU) + 64 + (K + 3) * max(0, args.iters_per_launch), force_num_latent_loss=FIT_OPTIONS["force_num_latent_loss"],
               halo_mode=args.halo_mode, **{k: v for k, v in model_opts.items() if k == "param_overrides"})
    data_host = dm.data_flattened
    f = fit.fused
    f.n_vox_global = W if cube else W * world
    W_total = W if cube else W * world
    n_state = f.n_state
    bytes_per_voxel = 8 * f.B + 24 * n_state          # data+tpts read; state/m/v read + written
    # algorithmic FP32 lane-instructions per voxel-iteration of each model family
    lane_instr_per_voxel = {"sim_art": 60 * 50 + 10 * 60 + 300, "real_like": 60 * 21 + 10 * 30 + 200,
                            "spatial": 60 * 21 + 10 * 30 + 200, "nn": 60 * 420 + 10 * 30 + 200,
                            "disp": 120000}[args.workload]
    # ... and algorithmic MUFU (XU-pipe) operations: 3 per element with the arterial term (1 without), 4 per
    # (voxel, sample), 4 per Box-Muller pair of draws; aslnn: 40 per (voxel, sample, time point) row (20 tanh)
    mufu_per_voxel = {"sim_art": 60 * 3 + 10 * 4 + 10 * 3 * 4, "real_like": 60 * 1 + 10 * 3 + 10 * 2 * 4,
                      "spatial": 60 * 1 + 10 * 3 + 10 * 2 * 4, "nn": 60 * 40 + 10 * 3 + 10 * 2 * 4,
                      "disp": None}[args.workload]

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
    f.finish()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[K])
    per_launch_ms = float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(K)]))
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = W_total * K / (total_ms * 1e-3)
    final_cost = float(f.cost_hist[f.step_count - 1].item()) / f.n_vox
    assert math.isfinite(final_cost), "non-finite cost"

    # ---- the same iterations, several per launch (svbasl_adam.n_iters): nothing is re-read between them ----
    fused = None
    kf = min(args.iters_per_launch, f.max_fuse)
    if kf > 1 and not f.mrf:
        for _ in range(3):
            f.step(kf)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            f.step(kf)
        e1.record()
        barrier()
        tf_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            td.all_reduce(tf_ms, op=td.ReduceOp.MAX)
        fused = {"iters_per_launch": kf, "launches": K, "value": W_total * K * kf / (float(tf_ms.item()) * 1e-3),
                 "unit": "voxel-iters/s", "ms_per_iteration": float(tf_ms.item()) / (K * kf),
                 "note": "same kernel, svbasl_adam.n_iters iterations per launch; not the headline value"}
        assert math.isfinite(float(f.cost_hist[f.step_count - 1].item()))

    # ---- end to end through the C ABI with HOST buffers (svb feeds each batch via feed_dict) ----
    from svb_models_asl_b200.ops import HostFeeder
    rows = list(range(0, f.T, f.n_batches))                           # the time points of batch 0
    numa_note = prefer_gpu_local_host_memory(local_rank)
    h_data = torch.from_numpy(np.ascontiguousarray(data_host.T[rows])).pin_memory()    # [B, ld]
    # time points in the low-rank form the model defines them by (t = ti + z*slicedt, aslrest.py:438-440): the
    # batch's B inversion times travel with every step, the per-voxel slice offset is resident like the mask
    lowrank = hasattr(model, "tpts_lowrank")
    if lowrank:
        ti_all, zoff = model.tpts_lowrank()
        h_ti = torch.from_numpy(np.ascontiguousarray(ti_all[rows])).pin_memory()
        zoff_dev = torch.as_tensor(zoff, device=dev) if zoff is not None else None
        h_tpts = None
    else:
        tp_full = np.broadcast_to(model.tpts(), data_host.shape)
        h_tpts = torch.from_numpy(np.ascontiguousarray(tp_full.T[rows])).pin_memory()
        h_ti = zoff_dev = None
    restore_host_placement()
    e2e_ok = not (f.mrf and world > 1 and args.halo_mode != "peer")   # host-fed spatial steps are one launch (peer mode)
    feeder = HostFeeder(f) if e2e_ok else None
    for i in range(WU if e2e_ok else 0):
        feeder.step(h_data, h_tpts, h_ti, zoff_dev)
    if e2e_ok:
        feeder.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(K if e2e_ok else 0):
        feeder.step(h_data, h_tpts, h_ti, zoff_dev)
    last_cost = feeder.sync() if e2e_ok else 0.0
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(te, op=td.ReduceOp.MAX)
    e2e_s = float(te.item())
    assert math.isfinite(last_cost)
    if feeder is not None:
        feeder.close()
    h2d = 4 * f.B * f.ld + (4 * f.B if lowrank else 4 * f.B * f.ld)

    f.release()
    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    peak, peak_src, _ = peaks()
    W = f.n_vox                                            # voxels one launch of this rank processes
    achieved = bytes_per_voxel * W / (per_launch_ms * 1e-3) / 1e9
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    fp32_peak = 148 * 128 * sm_hz
    line = {
        "metric": METRIC, "value": value, "unit": "voxel-iters/s", "n_gpus": world, "steps": K, "warmup": WU,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if cube else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"] + ", sample-based latent loss, Adam fused", "name": args.workload,
                   "voxels_per_gpu": W, "n_state": n_state,
                   **({"halo_mode": args.halo_mode} if (f.mrf and world > 1) else {}), "rng": "philox2x32-10 in-kernel",
                   "l2": "working set %.0f MB per step > 126 MB L2 (no flush needed)" % (bytes_per_voxel * W / 1e6)},
        "clocks": clocks,
        "fused": fused,
        "e2e": {"value": (W_total * K / e2e_s) if e2e_ok else None, "unit": "voxel-iters/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_s / K * 1e3,
                "h2d_gbs_implied": (h2d * K / e2e_s / 1e9) if e2e_ok else None, "host_memory": numa_note,
                "path": "svbasl_step_host: pinned host batch (data rows + the batch's TIs) -> H2D -> fused step -> "
                        "D2H cost, double-buffered"},
        # spatial iteration = pre-pass + step launch(es: interior + boundary slabs when sharded) + hyper step
        # a spatial iteration is ONE launch too (fused tail); the NCCL modes add the hyper-step launch
        "gpu_launches": K * (2 if (f.mrf and world > 1 and args.halo_mode != "peer") else 1),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(args.workload, W), "peak_source": peak_src,
                     "kernel": "step_kernel<%s, B=%d, %s>" % (wl["model"], f.B, "lean+spatial" if f.mrf else "lean"),
                     "algorithmic_bytes_per_voxel_iter": bytes_per_voxel, "avg_launch_ms": per_launch_ms,
                     "fp32": {"lane_instr_per_voxel_iter": lane_instr_per_voxel,
                              "achieved_tlane_per_s": lane_instr_per_voxel * W / (per_launch_ms * 1e-3) / 1e12,
                              "peak_tlane_per_s": fp32_peak / 1e12,
                              "frac": lane_instr_per_voxel * W / (per_launch_ms * 1e-3) / fp32_peak,
                              "note": "FP32-pipe roofline 148 SM x 128 lanes x measured SM clock (SURVEY 8d)"}},
        "final_mean_cost": final_cost,
    }
    if mufu_per_voxel:
        xu_peak = 148 * 16 * sm_hz                             # MUFU: 16 lanes per SM per clock
        xu_rate = mufu_per_voxel * W / (per_launch_ms * 1e-3)
        line["roofline"]["xu"] = {"mufu_per_voxel_iter": mufu_per_voxel, "achieved_tops_per_s": xu_rate / 1e12,
                                  "peak_tops_per_s": xu_peak / 1e12, "frac": xu_rate / xu_peak}
        fracs = {"hbm": line["roofline"]["frac"], "fp32": line["roofline"]["fp32"]["frac"],
                 "xu": line["roofline"]["xu"]["frac"]}
        line["roofline"]["binding"] = max(fracs, key=fracs.get)
    if wl["model"] == "aslnn":
        # Model.evaluate of the surrogate: FP32-pipe kernel vs tcgen05 tensor-core kernel, 200k voxels x S x B rows
        from svb_models_asl_b200.ops import evaluate_model, nn_evaluate_tc
        n_ev = 200_000
        pf = torch.rand(n_ev, 10, 1, device=dev) * 10 + 1
        pd = torch.rand(n_ev, 10, 1, device=dev) * 2 + 0.3
        tt = torch.as_tensor(np.asarray(model.tis, dtype=np.float32), device=dev).reshape(1, 1, -1)
        res = {}
        for name, fn in (("fp32_pipe", evaluate_model), ("tcgen05", lambda *a: nn_evaluate_tc(*a, check=False))):
            for _ in range(3):
                fn(model, [pf, pd], tt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(model, [pf, pd], tt)
            e1.record()
            torch.cuda.synchronize()
            res[name] = {"ms": e0.elapsed_time(e1) / 10, "rows_per_s": n_ev * 60 / (e0.elapsed_time(e1) / 10 * 1e-3)}
        line["nn_evaluate"] = res
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
