import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from oracle import asl_models as om
be = H.Backend("cuda")
cfg = om.AslConfig(tau=1.8, t1b=1.65, disp=True, casl=True)
rng = np.random.default_rng(5)
W = 3000
spec = H.aslrest_spec(cfg, n_samples=1, t_full=6)
prob = H.synth_problem(cfg, spec, W, rng, repeats=1, noise_sd=0.5)
n = spec.n_par
prob["state"][n:2*n] = -30.0          # no spread: theta = mu
prob["state"][2*n:] = 0.0
m = be.model_desc(cfg)
out = {}
for which in ("warp", "scalar"):
    if which == "scalar": os.environ["SVBASL_DISP_SCALAR"] = "1"
    e, bufs = be.engine_desc(spec, prob["state"], prob["data"], prob["tpts"], None, seed=5)
    cost, grad, csum = be.elbo_grad(m, e, spec.n_state, step=2)
    out[which] = (cost, grad)
    os.environ.pop("SVBASL_DISP_SCALAR", None)
cw, gw = out["warp"]; cs, gs = out["scalar"]
rel = np.abs(cw - cs) / np.maximum(np.abs(cs), 1e-6)
bad = np.where(rel > 1e-4)[0]
print("bad voxels", len(bad), "of", W)
names = cfg.param_names()
st = prob["state"]
def desc(w):
    delt = st[1, w]; s = np.exp(st[2, w]); sp = np.exp(st[3, w]); t0 = prob["tpts"][0, w]
    i0 = int(np.ceil(delt / 0.1 - 1e-9))
    return "delt %.4f s %.3f sp %.3f s*h %.3f u0 %.4f x1_0 %.4f tp0 %.3f" % (delt, s, sp, s * 0.1, i0 * 0.1 - delt, s * (i0 * 0.1 - delt), t0)
for w in bad[:25]:
    print("BAD ", w, "rel %.3e" % rel[w], desc(w), "cw %.4f cs %.4f" % (cw[w], cs[w]))
good = np.where(rel <= 1e-4)[0]
for w in good[:8]:
    print("good", w, desc(w))
sb = np.exp(st[2, bad]); sg = np.exp(st[2, good])
print("s range bad", sb.min() if len(bad) else None, sb.max() if len(bad) else None, " good", sg.min(), sg.max())
spb = np.exp(st[3, bad]); print("sp bad", np.sort(spb)[:10], np.sort(spb)[-10:] if len(bad) else None)
