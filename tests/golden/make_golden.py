#!/usr/bin/env python
"""
Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE SOURCE
(/root/reference/svb_models_asl/*.py) on seeded inputs.

TensorFlow / TFP / svb / fabber are not installable here, so the reference is
imported on top of (a) the numpy stand-ins in oracle/refshim (array backend only)
and (b) this repo's svb-compatible host layer (Model/ModelOption/get_parameter/
DataModel).  What is pinned is therefore the reference's *own Python code path*
(option resolution, parameter order, masks, formulas, composition) evaluated
with numpy/scipy arithmetic - not TensorFlow's kernels.

Run from the repo root in the build container (needs /root/reference):
    python tests/golden/make_golden.py
Inputs are float32-representable; each case stores the reference output computed
in float64 (`out64`) and in float32 (`out32`, the reference's own precision).
"""
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = ["/root/reference", os.path.join(ROOT, "oracle", "refshim"), ROOT]

import numpy as np  # noqa: E402

import svb  # noqa: E402  (this repo's alias package)
sys.modules.setdefault("svb.prior", types.ModuleType("svb.prior"))
import svb_models_asl as ref  # noqa: E402
from svb import DataModel  # noqa: E402

assert ref.__file__.startswith("/root/reference/"), ref.__file__

TIS = [2.05, 2.3, 2.55, 2.8, 3.05, 3.3]
PLDS = [0.25, 0.5, 0.75, 1.0, 1.25, 1.5]


def f32(x):
    return np.asarray(x, dtype=np.float32)


def run_eval(model, params32, t32):
    """params32 [P,W,S,1] float32, t32 [W,1,B] float32 -> (out64, out32)"""
    o64 = model.evaluate(list(params32.astype(np.float64)), t32.astype(np.float64))
    o32 = model.evaluate(list(params32), t32)
    return np.asarray(o64, dtype=np.float64), np.asarray(o32, dtype=np.float32)


def rand_params(rng, names, W, S):
    cols = []
    for n in names:
        if n in ("ftiss", "fwm"):
            v = rng.uniform(-2.0, 25.0, (W, S, 1))
        elif n in ("delttiss", "deltwm"):
            v = rng.uniform(0.05, 3.2, (W, S, 1))
        elif n in ("t1", "t1wm"):
            v = rng.uniform(0.8, 1.8, (W, S, 1))
        elif n == "fblood":
            v = rng.uniform(-1.0, 12.0, (W, S, 1))
        elif n == "deltblood":
            v = rng.uniform(-0.05, 2.6, (W, S, 1))
            v[::7] = rng.uniform(0.0005, 0.0099, v[::7].shape)      # below leadscale
        elif n == "s":
            v = rng.uniform(1.0, 20.0, (W, S, 1))
        elif n == "sp":
            v = rng.uniform(0.05, 12.0, (W, S, 1))                  # some above the clip at 10
        else:
            raise KeyError(n)
        cols.append(v)
    return f32(np.stack(cols, 0))


def voxel_tpts(rng, W, tis, repeats=1, slicedt=0.0452):
    z = rng.integers(0, 24, W)
    base = np.repeat(np.asarray(tis), repeats)
    return f32(base[None, :] + (z * slicedt)[:, None]).reshape(W, 1, -1)


def aslrest_cases():
    rng = np.random.default_rng(20260101)
    W, S = 48, 3
    out = {}
    cases = {
        "casl_tiss": dict(casl=True),
        "pasl_tiss": dict(casl=False),
        "casl_tiss_art": dict(casl=True, inferart=True),
        "pasl_tiss_art": dict(casl=False, inferart=True),
        "casl_noatt": dict(casl=True, inferatt=False),
        "casl_art_noatt": dict(casl=True, inferart=True, inferatt=False, _expect_error=True),
        "casl_artonly": dict(casl=True, artonly=True),
        "casl_t1": dict(casl=True, infert1=True),
        "pasl_t1_art": dict(casl=False, infert1=True, inferart=True),
        "casl_pvc": dict(casl=True, pvcorr=True, inferart=True,
                         pvgm=f32(rng.uniform(0.1, 0.6, W)), pvwm=f32(rng.uniform(0.0, 0.4, W))),
        "casl_pvc_t1": dict(casl=True, pvcorr=True, infert1=True,
                            pvgm=f32(rng.uniform(0.1, 0.6, W)), pvwm=f32(rng.uniform(0.0, 0.4, W))),
        # incwm without inferwm: the reference passes the float option `fwm` to tissue_signal,
        # which reads `.shape` from it (aslrest.py:292,328,352) -> AttributeError as shipped
        "casl_incwm_fixed": dict(casl=True, incwm=True, fwm=4.0, pvgm=0.6, pvwm=0.3, _expect_error=True),
        # inferwm without incwm / pvcorr: fwm and deltwm are parameters (aslrest.py:197-211) but no WM signal is
        # added (aslrest.py:327 tests incwm) and pc keeps its non-WM default 0.9 (aslrest.py:131-135)
        "casl_inferwm_noinc": dict(casl=True, inferwm=True, inferart=True),
    }
    for name, opts in cases.items():
        opts = dict(opts)
        expect_error = opts.pop("_expect_error", False)
        dm = DataModel(np.zeros((W, len(TIS)), dtype=np.float32))
        model = ref.AslRestModel(dm, tis=TIS, tau=1.8, t1b=1.65, repeats=1, **opts)
        names = [p.name for p in model.params]
        params = rand_params(rng, names, W, S)
        t = voxel_tpts(rng, W, TIS)
        rec = {"names": np.array(names), "params": params, "t": t,
               "prior_types": np.array([p.prior_type for p in model.params]),
               "prior_mean": f32([np.mean(p.prior_dist.mean) for p in model.params]),
               "prior_var": f32([np.mean(p.prior_dist.var) for p in model.params]),
               "post_mean": f32([np.mean(p.post_dist.mean) for p in model.params]),
               "post_var": f32([np.mean(p.post_dist.var) for p in model.params]),
               "pc": f32(np.mean(model.pc)), "attsd": f32(model.attsd), "artt": f32(model.artt)}
        for k in ("pvgm", "pvwm"):
            if k in opts:
                rec[k] = f32(opts[k])
        try:
            rec["out64"], rec["out32"] = run_eval(model, params, t)
            rec["error"] = np.array("")
        except Exception as exc:  # noqa: BLE001 - recording the reference's own failure
            if not expect_error:
                raise
            rec["error"] = np.array(type(exc).__name__)
        for k, v in rec.items():
            out["%s/%s" % (name, k)] = v
    return out


def aslrest_edges():
    """Boundary behaviour: t==delt, t==tau+delt, deltblood<leadscale, deltblood<=0."""
    dm = DataModel(np.zeros((1, 6), dtype=np.float32))
    out = {}
    model = ref.AslRestModel(dm, tis=TIS, tau=1.8, t1b=1.65, casl=True, inferart=True)
    t = f32([[0.001, 0.004, 0.006, 0.5, 0.7, 0.81, 1.0, 1.24, 1.245, 1.25, 1.255, 1.26, 1.7, 1.71,
              2.14, 2.5, 3.04, 3.045, 3.05, 3.055, 3.06, 3.3]]).reshape(1, 1, -1)
    deltb = f32([0.004, -0.1, 0.0, 0.01, 1.0, 1.2495, 0.0099999])
    delt = f32([0.7, 0.5, 1.25, 2.5, 1.24, 0.001, 3.3])
    W = deltb.size
    params = np.zeros((4, W, 1, 1), dtype=np.float32)
    params[0], params[1] = 10.0, delt.reshape(W, 1, 1)
    params[2], params[3] = 5.0, deltb.reshape(W, 1, 1)
    tt = np.repeat(t, W, axis=0)
    out["edges/params"], out["edges/t"] = params, tt
    out["edges/out64"], out["edges/out32"] = run_eval(model, params, tt)
    return out


def appendix_d():
    """The quick_test.py configuration (scripts/quick_test.py:10-33): ftiss 1,5,10, delt 1.3, t1b 1.6."""
    dm = DataModel(np.zeros((1, 6), dtype=np.float32))
    model = ref.AslRestModel(dm, tis=TIS, tau=1.8, t1b=1.6, t1=1.3, casl=True, repeats=1)
    params = np.zeros((2, 3, 1, 1), dtype=np.float32)
    params[0, :, 0, 0] = [1.0, 5.0, 10.0]
    params[1] = 1.3
    t = np.repeat(f32(TIS).reshape(1, 1, -1), 3, axis=0)
    o64, o32 = run_eval(model, params, t)
    return {"quick_test/params": params, "quick_test/t": t, "quick_test/out64": o64, "quick_test/out32": o32}


def tpts_and_init():
    """tpts() with slicedt on the real mask (aslrest.py:432-456) and the posterior initialisers
    (aslrest.py:461-520) on a slab of the real data."""
    sdir = "/root/reference/scripts"
    dm = DataModel(os.path.join(sdir, "asldata_diff.nii.gz"), mask=os.path.join(sdir, "asldata_mask.nii.gz"))
    model = ref.AslRestModel(dm, plds=PLDS, tau=1.8, casl=True, repeats=[8], slicedt=0.0452, inferart=True)
    t = model.tpts()
    sel = np.arange(0, dm.n_nodes, 97)
    data = dm.data_flattened
    out = {"real/n_nodes": np.array(dm.n_nodes), "real/tpts_sel": t[sel], "real/sel": sel,
           "real/tpts_sum": np.array(t.astype(np.float64).sum()), "real/data_sel": data[sel]}
    byname = {p.name: p for p in model.params}
    f, _ = model._init_flow(byname["ftiss"], t, data)
    fb, _ = model._init_fblood(byname["fblood"], t, data)
    d, dv = model._init_delt(byname["delttiss"], t, data)
    out["real/init_ftiss_sel"] = f32(f)[sel]
    out["real/init_fblood_sel"] = f32(fb)[sel]
    out["real/init_delt"] = f32(np.mean(d))
    out["real/init_delt_var"] = f32(np.mean(dv))
    model2 = ref.AslRestModel(dm, plds=PLDS, tau=1.8, casl=True, repeats=[8], slicedt=0.0452, att_init="max")
    d2, dv2 = model2._init_delt(model2.params[1], t, data)
    out["real/init_delt_max_sel"] = f32(d2)[sel]
    out["real/init_delt_max_var_sel"] = f32(dv2)[sel]
    return out


def disp_pieces():
    """AslRestDisp.evaluate cannot run as shipped; its building blocks can:
    aif_gammadisp (:69-110), resid_wellmix (:133-146), conv_tf (:148-171) and the tfp interpolation (:63)."""
    import tensorflow_probability as tfp
    rng = np.random.default_rng(7)
    dm = DataModel(np.zeros((1, 6), dtype=np.float32))
    model = ref.AslRestDisp(dm, tis=TIS, tau=1.8, t1b=1.65, casl=True, inferart=True)
    W, S = 6, 2
    delt = f32(rng.uniform(0.2, 2.0, (W, S, 1)))
    s = f32(rng.uniform(2.0, 15.0, (W, S, 1)))
    sp = f32(rng.uniform(0.1, 11.0, (W, S, 1)))
    grid = model.conv_t
    out = {"disp/names": np.array([p.name for p in model.params]), "disp/conv_t": grid,
           "disp/conv_nt": np.array(model.conv_nt), "disp/conv_tmax": np.array(model.conv_tmax),
           "disp/delt": delt, "disp/s": s, "disp/sp": sp,
           "disp/prior_mean": f32([np.mean(p.prior_dist.mean) for p in model.params]),
           "disp/prior_var": f32([np.mean(p.prior_dist.var) for p in model.params])}
    d64 = [x.astype(np.float64) for x in (delt, s, sp)]
    aif = np.asarray(model.aif_gammadisp(grid, d64[0], [d64[1], d64[2]]))
    out["disp/aif_as_written"] = aif
    resid = np.asarray(model.resid_wellmix(grid, 1.3))
    out["disp/resid"] = resid
    conv = np.asarray(model.conv_tf(aif, resid, model.conv_dt))
    out["disp/conv_of_aif"] = conv
    t = voxel_tpts(rng, W, TIS, slicedt=0.03).astype(np.float64)
    out["disp/t"] = t
    out["disp/interp"] = np.asarray(tfp.math.batch_interp_regular_1d_grid(t, 0, model.conv_tmax, conv, axis=-1))
    # a generic curve through conv_tf and the interpolation (independent of the gamma2-gamma2 defect)
    curve = rng.uniform(0, 2, (W, S, grid.size))
    out["disp/curve"] = curve
    out["disp/conv_of_curve"] = np.asarray(model.conv_tf(curve, resid, model.conv_dt))
    out["disp/interp_of_curve"] = np.asarray(
        tfp.math.batch_interp_regular_1d_grid(t, 0, model.conv_tmax, curve, axis=-1))
    # aif at the time points themselves (art_signal, :66-67)
    out["disp/aif_at_t_as_written"] = np.asarray(model.aif_gammadisp(t, d64[0], [d64[1], d64[2]]))
    # the non-dispersed test AIF (:112-131) needs PASL (tf.repeat on a [W,S,1] tensor mis-shapes under CASL)
    modelp = ref.AslRestDisp(dm, tis=TIS, tau=1.8, t1b=1.65, casl=False)
    out["disp/aif_nodisp_pasl"] = np.asarray(modelp.aif_nodisp(grid, d64[0], []))
    return out


def nn_case():
    """aslnn.evaluate (aslnn.py:93-126) with seeded weights in the reference .npy layout (:211-227,326-340)."""
    rng = np.random.default_rng(11)
    shapes = [(2, 10), (10, 10), (10, 1)]
    ws = [f32(rng.normal(0, 0.7, s)) for s in shapes]
    bs = [f32(rng.normal(0, 0.3, (1, s[1]))) for s in shapes]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for i, (w, b) in enumerate(zip(ws, bs)):
            np.save(os.path.join(d, "weights%i.npy" % i), w)
            np.save(os.path.join(d, "biases%i.npy" % i), b)
        dm = DataModel(np.zeros((1, 6), dtype=np.float32))
        model = ref.AslNNModel(dm, tis=TIS, tau=1.8, casl=True, train_load=d)
        W, S = 16, 4
        params = np.stack([f32(rng.uniform(0.5, 20, (W, S, 1))), f32(rng.uniform(0.1, 3.0, (W, S, 1)))], 0)
        t = voxel_tpts(rng, W, TIS, slicedt=0.0)
        o64 = np.asarray(model.evaluate(list(params.astype(np.float64)), t.astype(np.float64)))
        o32 = np.asarray(model.evaluate(list(params), t))
        out.update({"nn/names": np.array([p.name for p in model.params]),
                    "nn/dists": np.array([type(p.post_dist).__name__ for p in model.params]),
                    "nn/params": params, "nn/t": t, "nn/out64": o64, "nn/out32": f32(o32),
                    "nn/tpts_default": np.asarray(model.tpts())})
        for i, (w, b) in enumerate(zip(ws, bs)):
            out["nn/w%i" % i], out["nn/b%i" % i] = w, b
    return out


def main():
    blobs = {"aslrest_eval": aslrest_cases(), "aslrest_edges": {**aslrest_edges(), **appendix_d()},
             "aslrest_real": tpts_and_init(), "disp_pieces": disp_pieces(), "aslnn_eval": nn_case()}
    for name, blob in blobs.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **blob)
        print("%-16s %3i arrays  %7.1f KB" % (name, len(blob), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
