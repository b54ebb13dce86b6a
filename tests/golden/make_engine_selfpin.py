"""
Regression pin of the ENGINE oracle (oracle/svb_engine.py) against ITSELF.

The engine half of the oracle restates svb from recall and is not pinned to the reference (no svb source, no
TensorFlow in the image - "parity unpinned").  This fixture does not change that: it freezes the
oracle's own output on seeded inputs so that an accidental edit of the oracle (or of a torch upgrade changing its
arithmetic) shows up as a test failure instead of silently moving the target the kernels are compared with.

    python tests/golden/make_engine_selfpin.py        # rewrites tests/golden/engine_selfpin.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import asl_models as om          # noqa: E402
from tests import helpers as H               # noqa: E402

CASES = {
    "casl_tiss_art_numeric": (dict(casl=True, inferart=True), dict(latent="numeric")),
    "casl_tiss_art_analytic": (dict(casl=True, inferart=True), dict(latent="analytic")),
    "pasl_t1_llt": (dict(casl=False, infert1=True), dict(latent="analytic", cov="LLt")),
    "casl_spatial": (dict(casl=True), dict(mrf=(0,))),
    "disp": (dict(casl=True, inferart=True, disp=True), dict(n_samples=3)),
}


def compute(case):
    cfg_kw, spec_kw = CASES[case]
    rng = np.random.default_rng(sum(map(ord, case)))
    W = 12 if case != "casl_spatial" else 24
    cfg = om.AslConfig(tau=1.8, t1b=1.65, **cfg_kw)
    spec = H.aslrest_spec(cfg, **spec_kw)
    prob = H.synth_problem(cfg, spec, W, rng, noise_sd=0.5)
    eps = rng.normal(size=(spec.n_par, spec.n_samples, W)).astype(np.float32)
    kw = {}
    if case == "casl_spatial":
        from oracle import svb_engine as eng
        coords = np.stack(np.meshgrid(np.arange(2), np.arange(3), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
        kw = dict(hyper=np.asarray([-1.2]), neighbours=eng.neighbour_table(coords, (2, 3, 4)), grad_scale=1.0 / W)
    cost, grad, ghyper, _ = H.oracle_cost_grad(spec, prob, eps, **kw)
    out = {"cost": np.asarray(cost, dtype=np.float64), "grad": np.asarray(grad, dtype=np.float64)}
    if ghyper is not None:
        out["ghyper"] = np.asarray(ghyper, dtype=np.float64)
    return out


if __name__ == "__main__":
    rec = {}
    for case in CASES:
        for k, v in compute(case).items():
            rec["%s/%s" % (case, k)] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "engine_selfpin.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, {k: v.shape for k, v in rec.items()})
