import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from oracle import asl_models as om
be = H.Backend("cuda")
for kw in (dict(casl=True), dict(casl=True, inferart=True), dict(casl=False, inferart=True)):
    cfg = om.AslConfig(tau=1.8, t1b=1.65, disp=True, **kw)
    rng = np.random.default_rng(5)
    W = 4000
    spec = H.aslrest_spec(cfg, n_samples=10, t_full=6)
    prob = H.synth_problem(cfg, spec, W, rng, repeats=1, noise_sd=0.5)
    prob["state"][1] -= rng.uniform(0, 2.5, W) * (rng.uniform(size=W) < 0.2)     # some early / negative arrival times
    m = be.model_desc(cfg)
    out = {}
    for which in ("warp", "scalar"):
        if which == "scalar": os.environ["SVBASL_DISP_SCALAR"] = "1"
        e, bufs = be.engine_desc(spec, prob["state"], prob["data"], prob["tpts"], None, seed=5)
        out[which] = be.elbo_grad(m, e, spec.n_state, step=2)[:2]
        os.environ.pop("SVBASL_DISP_SCALAR", None)
    cw, gw = out["warp"]; cs, gs = out["scalar"]
    rel = np.abs(cw - cs) / np.maximum(np.abs(cs), 1e-6)
    grel = np.abs(gw - gs).max(axis=0) / np.maximum(np.abs(gs).max(axis=0), 1e-12)
    print(kw, "cost: max rel %.2e, n>1e-4: %d; grad per-voxel max rel %.2e n>1e-3: %d" % (rel.max(), (rel > 1e-4).sum(), grel.max(), (grel > 1e-3).sum()))
