#!/usr/bin/env python
"""
bench.py - headline benchmark: voxel-iterations/second of the fused ELBO+gradient(+Adam) step, S samples.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], scripts/asl_example_sim.py + gen_test_data.py): aslrest multi-PLD pCASL,
6 PLDs, ftiss + delttiss + arterial component (fblood ARD, deltblood), S = 10 samples, sample-based latent
loss, 1,000,000 synthetic voxels PER GPU (weak scaling; voxels are independent so shards need no data-path
collective).  A "step" is one launch of the fused kernel.  This workload has ONE time-point batch per epoch
(T = B = 6), so a launch fuses `--iters-per-launch` (16, the trainer's default) iterations: the state and the batch stay in registers and the
Adam moments in shared memory between them (svbasl_adam.n_iters; bit-identical to single-iteration launches,
tests/test_fit_gpu.py).  `value` = voxels x iterations / device time; the one-iteration-per-launch figure is reported
beside it (`single_launch`).

Additional legs on the same JSON line:
  sustained   the same launches repeated for >= 2 s with SM clocks sampled on every rank
  e2e         the same metric through svbasl_step_host with HOST buffers (one H2D batch copy per iteration)
  c5_strong   BASELINE.json configs[4]: ONE 215^3 (9.94 M voxel) aslrest volume with a spatial MRF prior on ftiss,
              6 PLD x 8 repeats, sharded over the N ranks (strong scaling; >= 2 s timed, clocks of every rank)
  shard_parity (N > 1) a 4,096-voxel spatial fit sharded over the N ranks against the same fit on one GPU, for all
              three halo modes
  cpu_baseline the oracle port on the host cores (bounded sample)
"""
import argparse
import ctypes as C
import json
import math
import os
import re
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
HEADLINE_DESC = "asl_example_sim: aslrest multi-PLD pCASL (6 PLD), ftiss+delttiss+arterial, S=10, B=T=6"
# scripts/asl_example.py:24-42 + spatial prior on ftiss (configs[4])
C5_MODEL_OPTIONS = {"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8], "slicedt": 0.0452,
                    "param_overrides": {"ftiss": {"prior_type": "M"}}}
C5_FIT = {"learning_rate": 0.01, "batch_size": 6, "sample_size": 10}

# The headline workload is BASELINE.json configs[1] ("sim_art").  The others are the remaining configs, kept
# for our own measurements (`--workload`); the driver only runs the default.
WORKLOADS = {
    "sim_art": dict(model="aslrest", options=MODEL_OPTIONS, batch=None, repeats=1, voxels=1_000_000, desc=HEADLINE_DESC),
    "real_like": dict(model="aslrest", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8], "slicedt": 0.0452},
                      batch=6, repeats=8, voxels=1_000_000,
                      desc="asl_example: aslrest 6 PLD x 8 repeats, slicedt, ftiss+delttiss, S=10, T=48, B=6"),
    "pvc": dict(model="aslrest", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "pvcorr": True,
                                          "pvgm": 0.6, "pvwm": 0.3, "inferart": True},
                batch=None, repeats=1, voxels=1_000_000,
                desc="aslrest PVEc (GM+WM tissue, aslrest.py:197-229) + arterial, P'=7, S=10, B=T=6"),
    "t1": dict(model="aslrest", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "infert1": True,
                                         "inferart": True},
               batch=None, repeats=1, voxels=1_000_000,
               desc="aslrest with infert1 (aslrest.py:221-229) + arterial, P'=6, S=10, B=T=6"),
    "disp": dict(model="aslrest_disp", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8], "inferart": True},
                 batch=6, repeats=8, voxels=250_000,
                 desc="aslrest_disp gamma dispersion, 6 PLD x 8 repeats, ftiss+delttiss+arterial+s+sp, S=10, T=48, B=6"),
    "nn": dict(model="aslnn", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1]}, batch=None, repeats=1,
               voxels=1_000_000, desc="aslnn MLP surrogate 2-10-10-1 (plugin defaults: 10x10 products of the fused step on tcgen05), "
                                      "ftiss+delttiss, S=10, B=T=6"),
    "nn_tc": dict(model="aslnn", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "tensor_core_step": True},
                  batch=None, repeats=1, voxels=1_000_000,
                  desc="aslnn MLP surrogate 2-10-10-1 with the 10x10 products of the fused step on tcgen05, S=10, B=T=6"),
    "nn_fp32": dict(model="aslnn", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "tensor_core_step": False},
                    batch=None, repeats=1, voxels=1_000_000,
                    desc="aslnn MLP surrogate 2-10-10-1, fused step on the FP32 pipe only, S=10, B=T=6"),
    "spatial": dict(model="aslrest", options={"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8],
                                              "param_overrides": {"ftiss": {"prior_type": "M"}}},
                    batch=6, repeats=8, voxels=1_000_000, cube=True,
                    desc="aslrest with spatial MRF prior on ftiss, 6 PLD x 8 repeats, S=10, T=48, B=6, 100^3 volume"),
}
# algorithmic FP32 lane-instructions and MUFU (XU-pipe) operations per voxel-iteration
# aslnn with the two 10x10 products per row on the tensor cores (the default, "nn" / "nn_tc"): the 200 FMAs per row leave
# the FP32 pipe, the 40 tanh MUFU pairs per row stay -> the XU pipe is the binding roofline of that kernel
LANE_INSTR = {"nn_tc": 60 * 220 + 10 * 30 + 200, "nn_fp32": 60 * 420 + 10 * 30 + 200, "sim_art": 60 * 50 + 10 * 60 + 300, "real_like": 60 * 21 + 10 * 30 + 200, "spatial": 60 * 21 + 10 * 30 + 200,
              "nn": 60 * 220 + 10 * 30 + 200, "disp": 120000, "pvc": 60 * 66 + 10 * 90 + 500, "t1": 60 * 60 + 10 * 75 + 400}
MUFU = {"nn_tc": 60 * 40 + 10 * 3 + 10 * 2 * 4, "nn_fp32": 60 * 40 + 10 * 3 + 10 * 2 * 4, "sim_art": 60 * 3 + 10 * 4 + 10 * 3 * 4, "real_like": 60 * 1 + 10 * 3 + 10 * 2 * 4, "spatial": 60 * 1 + 10 * 3 + 10 * 2 * 4,
        "nn": 60 * 40 + 10 * 3 + 10 * 2 * 4, "disp": None, "pvc": None, "t1": None}


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

    def __init__(self, gpu_index, period_ms=100):
        self.rows, self.proc, self.gpu, self.period = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def cpu_port_rate(n_vox, n_iters, seed=20260101, threads=None, chunk=250_000):
    """The oracle port of the reference graph (oracle/svb_engine.py: op-for-op, one materialised [W,S,B]
    tensor per elementary op, torch autograd backward, TF-form Adam), float32, all host threads.
    TensorFlow itself is not installable in this image.  Voxels are independent, so an iteration over `n_vox`
    voxels is run in chunks of `chunk` (bounds the host memory of the materialised intermediates; the gradient
    scale 1/n_vox of the mean cost is kept).  -> (voxel-iters/s, seconds per iteration, threads)"""
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
    gen_cfg = om.AslConfig(casl=True, inferart=True, tau=1.8, t1b=1.6)           # gen_test_data.py:28
    parts = []
    for c0 in range(0, n_vox, chunk):
        c1 = min(n_vox, c0 + chunk)
        t = torch.as_tensor(np.repeat(tis[:, None], c1 - c0, 1), dtype=torch.float32)
        par = [torch.as_tensor(truth[i, c0:c1]).reshape(c1 - c0, 1, 1) for i in range(4)]
        data = om.evaluate(gen_cfg, par, t.T.unsqueeze(1))[:, 0, :].T + torch.randn(6, c1 - c0)
        dnp = data.numpy()
        state = eng.initial_state(spec, [np.maximum(dnp.mean(0), 0.1), 1.3, np.maximum(dnp.max(0), 0.1), 1.3,
                                         np.log(np.maximum(dnp.var(0), 1.0))], [1.5, 1.0, 1.5, 1.0, 1.02], c1 - c0,
                                  dtype=torch.float32)
        parts.append({"t": t, "data": data, "state": state, "opt": eng.Adam(lr=0.05)})
    hyper = torch.zeros(0)
    times = []
    warm = 1 if n_vox >= 500_000 else 2       # thread pool / allocator warm-up iterations, not timed
    for it in range(n_iters + warm):
        t0 = time.perf_counter()
        for p in parts:
            eps = torch.randn(5, 10, p["data"].shape[1])
            _cost, gs, _gh, _ = eng.cost_and_grad(spec, p["state"], hyper, p["data"], p["t"], eps, grad_scale=1.0 / n_vox)
            p["opt"].update({"state": (p["state"], gs)})
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times[warm:]))
    return n_vox / sec, sec, threads


def run_reference(args, out):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, same metric and config:
    every step is one iteration over the SAME 1,000,000 voxels the GPU arm processes per launch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_vox = args.voxels or WORKLOADS["sim_art"]["voxels"]
    steps = max(1, min(args.steps, args.reference_max_steps))
    rate, sec, threads = cpu_port_rate(n_vox, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "voxel-iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": HEADLINE_DESC + ", sample-based latent loss, Adam fused", "name": "sim_art",
                   "voxels_per_gpu": n_vox, "timed_iterations": steps,
                   "note": "one step = one iteration over all %d voxels (in chunks of 250k to bound host memory); the "
                           "median of %d timed iterations is reported" % (n_vox, steps)},
        "cpu_baseline": {"value": rate, "unit": "voxel-iters/s", "cores": threads, "kind": "port",
                         "sample": "%d voxels x %d iterations (%.2f s/iter); PyTorch-CPU op-for-op restatement of the "
                                   "reference TF graph (TensorFlow/svb not installable in this image)" % (n_vox, steps, sec)},
        "e2e": {"value": rate, "unit": "voxel-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()


# ----------------------------------------------------------------------------------------------------
def _cpus_of_list(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            if lo.strip().isdigit():
                cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _gpu_locality(dev_index):
    """(numa node or None, set of local cpus, how) for the GPU: sysfs first, then `nvidia-smi topo -m`."""
    try:
        import torch
        props = torch.cuda.get_device_properties(dev_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        if node >= 0:
            return node, _cpus_of_list(open(base + "/local_cpulist").read()), "sysfs %s" % bdf
    except Exception:                                                              # noqa: BLE001
        pass
    try:
        # physical index of this process's device (CUDA_VISIBLE_DEVICES may renumber)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[dev_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else dev_index
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        head = None
        for line in topo.splitlines():
            clean = re.sub(r"\x1b\[[0-9;]*m", "", line)
            cols = re.split(r"\t+|\s{2,}", clean.strip())
            if "CPU Affinity" in clean:
                head = cols
                continue
            if head and cols and cols[0] == "GPU%d" % phys:
                # data rows have one leading label column more than the header
                off = len(cols) - len(head)
                ia = head.index("CPU Affinity") + off
                cpus = _cpus_of_list(cols[ia]) if ia < len(cols) else set()
                node = None
                if "NUMA Affinity" in head:
                    inu = head.index("NUMA Affinity") + off
                    if inu < len(cols) and cols[inu].split(",")[0].strip().isdigit():
                        node = int(cols[inu].split(",")[0])
                if node is None and cpus:
                    for nd in sorted(os.listdir("/sys/devices/system/node")):
                        if nd.startswith("node") and cpus & _cpus_of_list(open("/sys/devices/system/node/%s/cpulist" % nd).read()):
                            node = int(nd[4:])
                            break
                return node, cpus, "nvidia-smi topo GPU%d" % phys
    except Exception:                                                              # noqa: BLE001
        pass
    return None, set(), "unknown"


def prefer_gpu_local_host_memory(dev_index):
    """Best effort, Linux only: ask the kernel to place this process's NEW host pages (the pinned staging buffers
    of the end-to-end leg) on the NUMA node the GPU hangs off, and run on that node's CPUs when the cpuset allows.
    Returns a short description for the JSON line.  Never fatal.  Undone by restore_host_placement() once the
    buffers exist (the CPU baseline must see every core)."""
    global _SAVED_AFFINITY
    _SAVED_AFFINITY = os.sched_getaffinity(0)
    try:
        node, cpus, how = _gpu_locality(dev_index)
        if node is None and not cpus:
            return "gpu numa node unknown (sysfs and nvidia-smi topo gave nothing)"
        note = "gpu on numa node %s (%s)" % (node, how)
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
            note += ", cpus pinned to %d local" % len(cpus & allowed)
        else:
            note += ", no local cpu in cpuset"
        if node is not None:
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


class Dist:
    """torch.distributed plumbing of the bench (barrier, max-over-ranks, gather of small objects)."""

    def __init__(self, torch, td, dev, world):
        self.torch, self.td, self.dev, self.world = torch, td, dev, world

    def barrier(self):
        if self.world > 1:
            self.td.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.td.all_gather_object(out, obj)
        return out


def timed_steps(D, torch, step_fn, n):
    """n calls of step_fn bracketed by barrier + synchronize, CUDA events on the launch stream -> max over ranks (ms),
    mean duration of one call on this rank (ms)."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    D.barrier()
    ev[0].record()
    for i in range(n):
        step_fn()
        ev[i + 1].record()
    D.barrier()
    total = ev[0].elapsed_time(ev[n])
    per = float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(n)]))
    return D.max(total), per


def sustained_leg(D, torch, step_fn, est_ms_per_call, local_rank, seconds=2.0, max_calls=200000):
    """step_fn repeated for at least `seconds` (same count on every rank), SM clocks sampled on every rank."""
    n = int(min(max_calls, max(10, math.ceil(seconds * 1.15 * 1e3 / max(est_ms_per_call, 1e-3)))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    sampler = ClockSampler(local_rank).start()
    time.sleep(0.12)
    e0.record()
    for _ in range(n):
        step_fn()
    e1.record()
    D.barrier()
    clocks = sampler.stop()
    ms = D.max(e0.elapsed_time(e1))
    return n, ms, D.gather(clocks)


# ----------------------------------------------------------------------------------------------------
def c5_volume_shard(torch, dev, side, plan, halo, n_t_rep=8, chunk=1 << 20):
    """Synthetic multi-PLD data of the local voxel range [plan.lo - halo_lo, plan.hi + halo_hi) of a side^3 volume,
    deterministic per GLOBAL voxel (chunks of 2^20 voxels of the global order are generated from their own seeds, a
    rank generates the chunks that overlap its range): gen_test_data.py restated, tissue only, noise sd 1.
    -> data [T, ld] on the device, zoff [ld] (slice offset z*slicedt)."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    T = len(PLDS) * n_t_rep
    gen = AslRestModel(DataModel(np.zeros((1, T), dtype=np.float32)), tau=1.8, casl=True, plds=PLDS, repeats=[n_t_rep], t1b=1.6)
    tis = np.repeat(np.asarray(gen.tis, dtype=np.float32), n_t_rep)
    g0, g1 = plan.lo - halo[0], plan.hi + halo[1]
    data = torch.empty((T, g1 - g0), device=dev, dtype=torch.float32)
    for c in range(g0 // chunk, (g1 - 1) // chunk + 1):
        a, b = c * chunk, min((c + 1) * chunk, side ** 3)
        rng = np.random.default_rng(7_000_000 + c)
        ftiss = rng.uniform(1.0, 20.0, b - a).astype(np.float32)
        delt = rng.uniform(0.6, 2.5, b - a).astype(np.float32)
        z = (np.arange(a, b) % side).astype(np.float32) * 0.0452
        t = torch.as_tensor(tis[None, :] + z[:, None], device=dev).reshape(b - a, 1, T)
        sig = gen.evaluate([ftiss.reshape(-1, 1, 1), delt.reshape(-1, 1, 1)], t)[:, 0, :]
        sig = sig + torch.randn(sig.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(9_000_000 + c))
        lo, hi = max(a, g0), min(b, g1)
        data[:, lo - g0:hi - g0] = sig[lo - a:hi - a].T
        del sig, t
    zoff = torch.as_tensor(((np.arange(g0, g1) % side) * 0.0452).astype(np.float32), device=dev)
    return data, tis, zoff


def build_spatial_fit(torch, dev, side, rank, world, halo_mode, max_steps, use_graph=True):
    """SvbFit of the aslrest model with a spatial MRF prior on ftiss on a side^3 volume sharded over `world` ranks,
    data generated on the device (SvbFit.setup_from_device)."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.sharding import ShardPlan
    from svb_models_asl_b200.svbcompat.fit import SvbFit
    dm = DataModel.header((side, side, side), len(PLDS) * 8)
    model = AslRestModel(dm, **C5_MODEL_OPTIONS)
    fit = SvbFit(dm, model)
    fit.rank, fit.world = rank, world
    plan = ShardPlan(dm.n_nodes, rank, world, dm.neighbour_table)
    data, tis, zoff = c5_volume_shard(torch, dev, side, plan, (plan.halo_lo, plan.halo_hi))
    f = fit.setup_from_device(data, plan, tis, zoff, C5_FIT["batch_size"], C5_FIT["sample_size"], C5_FIT["learning_rate"],
                              max_steps, halo_mode=halo_mode, use_graph=use_graph,
                              param_overrides=C5_MODEL_OPTIONS["param_overrides"])
    return fit, f, plan


def c5_strong_leg(D, torch, dev, rank, world, local_rank, args):
    side = args.c5_side
    fit, f, plan = build_spatial_fit(torch, dev, side, rank, world, args.halo_mode, max_steps=args.c5_max_iters + 400)
    for _ in range(30):
        f.step()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f.step()
    e1.record()
    D.barrier()
    est = D.max(e0.elapsed_time(e1)) / 20
    n, ms, clocks = sustained_leg(D, torch, f.step, est, local_rank, seconds=args.c5_seconds, max_calls=args.c5_max_iters)
    f.check_peers()
    cost = float(f.cost_hist[f.step_count - 1].item())
    tot = torch.tensor([cost], device=dev, dtype=torch.float64)
    if world > 1:
        D.td.all_reduce(tot)
    log_ak = float(f.log_ak[0].item())
    n_launch = {"prepass": 3, "separate_tail": 2, "fused": 1}[f.spatial_flow] + (1 if (world > 1 and args.halo_mode != "peer") else 0)
    f.release()
    W = side ** 3
    res = {"workload": "asl_example: aslrest 6 PLD x 8 repeats, slicedt, spatial MRF prior on ftiss, S=10, T=48, B=6, "
                       "ONE %d^3 volume sharded over the ranks (x-slabs)" % side,
           "voxels": W, "n_gpus": world, "halo_mode": args.halo_mode if world > 1 else None, "flow": f.spatial_flow,
           "iters": n, "seconds": ms * 1e-3, "ms_per_iter": ms / n, "value": W * n / (ms * 1e-3), "unit": "voxel-iters/s",
           "scaling": "strong", "launches_per_iter": n_launch, "clocks_per_rank": clocks,
           "final_mean_cost": float(tot.item()) / W, "log_ak": log_ak,
           "halo_voxels_per_rank": plan.halo_lo + plan.halo_hi,
           "nvlink_bytes_per_iter_per_rank": 4 * f.S * len(f.mrf) * (plan.prev_halo_hi + plan.next_halo_lo)}
    assert math.isfinite(res["final_mean_cost"]) and math.isfinite(log_ak)
    del fit, f
    torch.cuda.empty_cache()
    return res


def shard_parity_leg(D, torch, dev, rank, world, args):
    """A 16^3 spatial fit sharded over all ranks, every halo mode, against the same fit on ONE GPU (rank 0):
    max relative difference of the posterior state rows and of log ak after `n_it` iterations."""
    side, n_it = 16, 40
    out = {"voxels": side ** 3, "iterations": n_it}
    ref_state = ref_lak = None
    if rank == 0:
        fit1, f1, _p = build_spatial_fit(torch, dev, side, 0, 1, "peer", max_steps=n_it + 8)
        for _ in range(n_it):
            f1.step()
        torch.cuda.synchronize()
        ref_state = f1.state.cpu().numpy()
        ref_lak = float(f1.log_ak[0].item())
        f1.release()
    for mode in ("peer", "peer+nccl", "nccl"):
        fit, f, plan = build_spatial_fit(torch, dev, side, rank, world, mode, max_steps=n_it + 8)
        for _ in range(n_it):
            f.step()
        f.check_peers()
        torch.cuda.synchronize()
        own = f.state[:, plan.halo_lo:plan.halo_lo + plan.n_own].cpu().numpy()
        parts = D.gather(own)
        lak = float(f.log_ak[0].item())
        f.release()
        if rank == 0:
            st = np.concatenate(parts, axis=1)
            scale = np.maximum(np.abs(ref_state).max(axis=1, keepdims=True), 1e-30)
            out[mode] = {"max_rel_state": float((np.abs(st - ref_state) / scale).max()),
                         "log_ak_rel": abs(lak - ref_lak) / max(abs(ref_lak), 1e-30)}
        del fit, f
    return out


# ----------------------------------------------------------------------------------------------------
def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--voxels", type=int, default=None, help="voxels per GPU")
    ap.add_argument("--workload", default="sim_art", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-voxels", type=int, default=100_000, help="voxels of the bounded CPU sample (cpu_baseline)")
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--reference-max-steps", type=int, default=8,
                    help="--impl reference: at most this many timed iterations over the full 1M voxels (~1-2 s each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--iters-per-launch", type=int, default=16,
                    help="iterations fused per launch when the workload has one time-point batch per epoch (state, "
                         "data and Adam moments stay on chip between them); 1 = one iteration per launch")
    ap.add_argument("--halo-mode", default="peer", choices=["peer", "peer+nccl", "nccl"],
                    help="spatial prior on N > 1 GPUs: how halo samples and the log-ak gradient travel")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-c5", action="store_true", help="skip the 10M-voxel spatial strong-scaling leg")
    ap.add_argument("--c5-side", type=int, default=215)
    ap.add_argument("--c5-seconds", type=float, default=2.0)
    ap.add_argument("--c5-max-iters", type=int, default=20000)
    ap.add_argument("--no-shard-parity", action="store_true")
    ap.add_argument("--only-c5", action="store_true", help="run the c5_strong (+ shard_parity) leg alone (our own sweeps)")
    ap.add_argument("--sweep-env", default=None,
                    help="our own tuning sweeps: 'NAME=v1,v2;NAME2=w1,w2' - times the workload's launches under every "
                         "combination of these environment switches of the library (read per launch), prints a table to "
                         "stderr and exits")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import torch.distributed as td
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.ops import HostFeeder
    from svb_models_asl_b200.plugin import get_model_class
    from svb_models_asl_b200.svbcompat.fit import SvbFit

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    D = Dist(torch, td, dev, world)
    if args.only_c5:
        c5 = c5_strong_leg(D, torch, dev, rank, world, local_rank, args)
        parity = shard_parity_leg(D, torch, dev, rank, world, args) if (world > 1 and not args.no_shard_parity) else None
        if rank == 0:
            out.write(json.dumps({"c5_strong": c5, "shard_parity": parity}) + "\n")
            out.flush()
        if world > 1:
            td.destroy_process_group()
        return
    wl = WORKLOADS[args.workload]
    W, K, WU = args.voxels or wl["voxels"], args.steps, max(3, args.warmup)
    cube = bool(wl.get("cube"))            # one shared volume, sharded over the ranks (strong scaling)
    if cube:
        side = int(round(W ** (1.0 / 3.0)))
        W = side ** 3

    # ---- synthetic shard (gen_test_data.py restated; generated through the plugin's own evaluate kernel) ----
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
    if cube:
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
    if not cube:
        fit.lo, fit.hi = 0, W                                           # every rank owns its own W voxels (weak scaling)
    n_t = len(PLDS) * reps
    n_batches = int(math.ceil(n_t / (wl["batch"] or n_t)))
    ipl = max(1, min(args.iters_per_launch, 64)) if not cube else 1     # spatial priors: one iteration per launch
    est_iters = (3 * (K + WU) + 64) * ipl + 200_000
    fit._setup(model.tpts(), dm.data_flattened, wl["batch"], FIT_OPTIONS["sample_size"], FIT_OPTIONS["learning_rate"],
               epochs=est_iters, force_num_latent_loss=FIT_OPTIONS["force_num_latent_loss"],
               halo_mode=args.halo_mode, **{k: v for k, v in model_opts.items() if k == "param_overrides"})
    data_host = dm.data_flattened
    f = fit.fused
    f.n_vox_global = W if cube else W * world
    W_total = W if cube else W * world
    n_state = f.n_state
    # data+tpts read (one batch per launch when every iteration sees the same batch, else one per fused iteration);
    # state/m/v read + written once per launch
    bytes_per_voxel = 8 * f.B * (1 if n_batches == 1 else ipl) + 24 * n_state
    lane_instr = LANE_INSTR[args.workload]
    mufu = MUFU[args.workload]

    # ---- device-resident timing: K launches of ipl iterations, CUDA events on the launch stream ----
    step_fn = (lambda: f.step(ipl)) if not f.mrf else f.step
    for _ in range(WU):
        step_fn()
    if args.sweep_env:
        import itertools
        axes = [(kv.split("=")[0], kv.split("=")[1].split(",")) for kv in args.sweep_env.split(";") if kv]
        for combo in itertools.product(*[v for _k, v in axes]):
            for (k, _v), val in zip(axes, combo):
                os.environ[k] = val
            for _ in range(2):
                step_fn()
            ms, _per = timed_steps(D, torch, step_fn, K)
            if rank == 0:
                sys.stderr.write("sweep %s: %.4f ms per launch, %.4g voxel-iters/s\n"
                                 % (" ".join("%s=%s" % (k, v) for (k, _), v in zip(axes, combo)), ms / K,
                                    W_total * K * ipl / (ms * 1e-3)))
        return
    sampler = ClockSampler(local_rank).start()
    total_ms, per_launch_ms = timed_steps(D, torch, step_fn, K)
    clocks_k = sampler.stop()
    value = W_total * K * ipl / (total_ms * 1e-3)
    final_cost = float(f.cost_hist[f.step_count - 1].item()) / f.n_vox
    assert math.isfinite(final_cost), "non-finite cost"

    # ---- one iteration per launch (round-1 headline), for continuity ----
    single = None
    if ipl > 1:
        for _ in range(3):
            f.step(1)
        s_ms, s_per = timed_steps(D, torch, lambda: f.step(1), K)
        single = {"iters_per_launch": 1, "launches": K, "value": W_total * K / (s_ms * 1e-3), "unit": "voxel-iters/s",
                  "ms_per_iteration": s_ms / K, "avg_launch_ms": s_per,
                  "hbm_gbs": (8 * f.B + 24 * n_state) * f.n_vox / (s_per * 1e-3) / 1e9}

    # ---- sustained: the headline launches for >= 2 s, SM clocks of every rank ----
    n_sus, sus_ms, sus_clocks = sustained_leg(D, torch, step_fn, per_launch_ms, local_rank, seconds=args.sustained_seconds,
                                              max_calls=(f.lr_t.numel() - f.step_count - (2 * (K + WU) + 64) * ipl) // ipl)
    sustained = {"launches": n_sus, "seconds": sus_ms * 1e-3, "value": W_total * n_sus * ipl / (sus_ms * 1e-3),
                 "unit": "voxel-iters/s", "clocks_per_rank": sus_clocks}
    assert math.isfinite(float(f.cost_hist[f.step_count - 1].item()))

    # ---- end to end through the C ABI with HOST buffers (svb feeds each batch via feed_dict) ----
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
    e2e_ok = e2e_ok and not (f.mrf and f.graphs is not None and world > 1)
    feeder = None
    if e2e_ok and f.mrf and f.halo[0] + f.halo[1] > 0:
        # the host batch of a sharded volume must cover the local range (halo columns unused by the step)
        h_full = np.zeros((len(rows), f.ld), dtype=np.float32)
        h_full[:, f.halo[0]:f.halo[0] + f.n_vox] = data_host.T[rows][:, fit.lo:fit.hi]
        h_data = torch.from_numpy(h_full).pin_memory()
    if e2e_ok:
        if f.mrf:
            f.graphs = None                    # host-fed iterations are launched directly (svbasl_step_host)
        feeder = HostFeeder(f)
        for i in range(WU):
            feeder.step(h_data, h_tpts, h_ti, zoff_dev)
        feeder.sync()
    D.barrier()
    t0 = time.perf_counter()
    for i in range(K if e2e_ok else 0):
        feeder.step(h_data, h_tpts, h_ti, zoff_dev)
    last_cost = feeder.sync() if e2e_ok else 0.0
    e2e_s = D.max(max(time.perf_counter() - t0, 1e-9))
    assert math.isfinite(last_cost)
    if feeder is not None:
        feeder.close()
    h2d = 4 * f.B * f.ld + (4 * f.B if lowrank else 4 * f.B * f.ld)
    n_vox_launch = f.n_vox
    spatial_flow = getattr(f, "spatial_flow", None) if f.mrf else None
    f.release()
    del fit, f, feeder
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs[4]: one 10M-voxel volume, spatial prior, sharded over the ranks ----
    c5 = parity = None
    if args.workload == "sim_art" and not args.no_c5:
        c5 = c5_strong_leg(D, torch, dev, rank, world, local_rank, args)
        if world > 1 and not args.no_shard_parity:
            parity = shard_parity_leg(D, torch, dev, rank, world, args)

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    peak, peak_src, _ = peaks()
    clk = [c.get("sm_mhz") for c in sus_clocks if c.get("sm_mhz")]
    sm_mhz = float(np.median(clk)) if clk else (clocks_k.get("sm_mhz") or 1965.0)
    fp32_peak = 148 * 128 * sm_mhz * 1e6
    iter_rate = n_vox_launch * ipl / (per_launch_ms * 1e-3)             # voxel-iterations per second of one launch
    hbm_ach = bytes_per_voxel * n_vox_launch / (per_launch_ms * 1e-3) / 1e9
    fracs = {"hbm": hbm_ach / peak, "fp32": lane_instr * iter_rate / fp32_peak}
    roof = {"kernel": "step_kernel<%s, B=%d, %s>" % (wl["model"], WORKLOADS[args.workload]["batch"] or n_t,
                                                     "lean+spatial" if spatial_flow else "lean"),
            "iterations_per_launch": ipl, "avg_launch_ms": per_launch_ms,
            "traffic": measured_traffic(args.workload, n_vox_launch), "peak_source": peak_src,
            "hbm": {"achieved": hbm_ach, "peak": peak, "unit": "GB/s", "frac": fracs["hbm"],
                    "algorithmic_bytes_per_voxel_per_launch": bytes_per_voxel},
            "fp32": {"lane_instr_per_voxel_iter": lane_instr, "achieved": lane_instr * iter_rate / 1e12,
                     "peak": fp32_peak / 1e12, "unit": "T lane-instr/s", "frac": fracs["fp32"],
                     "note": "FP32-pipe roofline 148 SM x 128 lanes x the median SM clock of the sustained leg "
                             "(%.0f MHz); algorithmic instruction count of SURVEY 8d" % sm_mhz}}
    if mufu:
        xu_peak = 148 * 16 * sm_mhz * 1e6                              # MUFU: 16 lanes per SM per clock
        fracs["xu"] = mufu * iter_rate / xu_peak
        roof["xu"] = {"mufu_per_voxel_iter": mufu, "achieved": mufu * iter_rate / 1e12, "peak": xu_peak / 1e12,
                      "unit": "T op/s", "frac": fracs["xu"]}
    binding = max(fracs, key=fracs.get)
    # the headline fraction is the BINDING roofline's (the largest of the three; SURVEY 8d names it per model family)
    roof.update({"bound": binding, "binding": binding, "achieved": roof[binding]["achieved"], "peak": roof[binding]["peak"],
                 "unit": roof[binding]["unit"], "frac": fracs[binding]})
    line = {
        "metric": METRIC, "value": value, "unit": "voxel-iters/s", "n_gpus": world, "steps": K, "warmup": WU,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if cube else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"] + ", sample-based latent loss, Adam fused", "name": args.workload,
                   "voxels_per_gpu": n_vox_launch, "n_state": n_state, "iters_per_launch": ipl,
                   "step": "one launch = %d iteration(s) over the shard's voxels" % ipl,
                   **({"halo_mode": args.halo_mode, "flow": spatial_flow} if (spatial_flow and world > 1) else {}),
                   "rng": "philox2x32-10 in-kernel",
                   "l2": "working set %.0f MB per launch > 126 MB L2 (no flush needed)" % (bytes_per_voxel * n_vox_launch / 1e6)},
        "clocks": {**clocks_k, "sustained_sm_mhz_median": sm_mhz,
                   "sustained_reasons": sorted({r for c in sus_clocks for r in (c.get("reasons") or [])}),
                   "sustained_samples_per_rank": [c.get("samples") for c in sus_clocks]},
        "single_launch": single,
        "sustained": sustained,
        "e2e": {"value": (W_total * K / e2e_s) if e2e_ok else None, "unit": "voxel-iters/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_s / K * 1e3, "iterations_per_step": 1,
                "h2d_gbs_implied": (h2d * K / e2e_s / 1e9) if e2e_ok else None, "host_memory": numa_note,
                "path": "svbasl_step_host: pinned host batch (data rows + the batch's TIs) -> H2D -> fused step -> "
                        "D2H cost, double-buffered; ONE iteration per host batch, as the reference feeds every "
                        "sess.run (no multi-iteration fusing here)"},
        "gpu_launches": K * ({"prepass": 3, "separate_tail": 2, "fused": 1}.get(spatial_flow, 1)),
        "roofline": roof,
        "final_mean_cost": final_cost,
    }
    if c5 is not None:
        line["c5_strong"] = c5
    if parity is not None:
        line["shard_parity"] = parity
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
