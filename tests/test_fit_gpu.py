"""
End-to-end fits through the reference-facing API (svb.main.run / SvbFit) on the GPU.

* the option dicts of scripts/asl_example_sim.py and scripts/asl_example.py drive the engine unchanged
  (minus plotting), outputs have the reference's names and shapes (asl_example.py:47-54);
* converged posterior means agree with the oracle's own fit run on IDENTICAL draws (the Philox stream is
  written out with svbasl_fill_eps and fed to the oracle) within BASELINE.json's 1e-3 relative;
* multi-iteration launches (iters_per_launch) are bit-compatible with single-iteration launches.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import asl_models as om
from oracle import svb_engine as eng
from tests import helpers as H

pytestmark = pytest.mark.gpu

PLDS = [0.25, 0.5, 0.75, 1.0, 1.25, 1.5]


def _sim_volume(shape, rng, repeats=1, noise=0.5, t1b=1.6, slicedt=0.0):
    """gen_test_data.py restated (ftiss~U(1,20), delttiss~U(0.6,2.5), generator t1b=1.6)."""
    n = int(np.prod(shape))
    ftiss = rng.uniform(1.0, 20.0, n)
    delt = rng.uniform(0.6, 2.5, n)
    cfg = om.AslConfig(casl=True, tau=1.8, t1b=t1b)
    tis = np.repeat(np.asarray([1.8 + p for p in PLDS]), repeats)
    z = np.arange(n) % shape[2]
    t = tis[None, :] + (z * slicedt)[:, None]
    sig = om.evaluate(cfg, [torch.as_tensor(ftiss).reshape(n, 1, 1), torch.as_tensor(delt).reshape(n, 1, 1)],
                      torch.as_tensor(t).unsqueeze(1))[:, 0, :].numpy()
    sig = sig + rng.normal(0, noise, sig.shape)
    return sig.reshape(*shape, -1).astype(np.float32), ftiss.reshape(shape), delt.reshape(shape)


def test_asl_example_sim_options_run_and_recover_truth(tmp_path):
    from svb.main import run
    from svb_models_asl_b200.svbcompat import nifti
    rng = np.random.default_rng(1)
    vol, ftiss, delt = _sim_volume((10, 10, 10), rng, noise=0.2)
    nifti.save(vol, str(tmp_path / "sig.nii.gz"))
    options = {            # scripts/asl_example_sim.py:23-40, epochs shortened
        "tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "learning_rate": 0.05, "sample_size": 10,
        "epochs": 1500, "log_stream": None, "save_mean": True, "save_var": True, "save_param_history": True,
        "save_cost": True, "save_cost_history": True, "save_model_fit": True, "save_log": True,
        "force_num_latent_loss": True, "display_step": 0,
    }
    out = str(tmp_path / "out")
    runtime, svb, history = run(str(tmp_path / "sig.nii.gz"), "aslrest", out, **options)
    for name in ("mean_ftiss", "mean_delttiss", "var_ftiss", "var_delttiss", "cost", "cost_history", "modelfit",
                 "mean_ftiss_history", "mean_delttiss_history"):
        assert os.path.exists(os.path.join(out, name + ".nii.gz")), name
    assert os.path.exists(os.path.join(out, "logfile"))
    mf = nifti.load(os.path.join(out, "mean_ftiss.nii.gz")).data
    md = nifti.load(os.path.join(out, "mean_delttiss.nii.gz")).data
    assert mf.shape == (10, 10, 10) and md.shape == (10, 10, 10)
    assert nifti.load(os.path.join(out, "modelfit.nii.gz")).data.shape == (10, 10, 10, 6)
    assert nifti.load(os.path.join(out, "cost_history.nii.gz")).data.shape == (10, 10, 10, 1501)
    # t1b mismatch between generator (1.6) and fit (1.65) is the reference's own; recovery is still close
    assert np.median(np.abs(mf - ftiss) / ftiss) < 0.06
    assert np.median(np.abs(md - delt)) < 0.08
    assert history["mean_cost"][-1] < history["mean_cost"][0]
    assert history["nan_skips"] == 0 and runtime > 0


def test_converged_posterior_matches_oracle_fit_on_identical_draws():
    """BASELINE.json: converged posterior means of ftiss/delttiss within 1e-3 relative."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.svbcompat.fit import SvbFit
    rng = np.random.default_rng(2)
    W = 256
    vol, _f, _d = _sim_volume((4, 8, 8), rng, noise=0.3, t1b=1.65)
    data = vol.reshape(W, -1)
    dm = DataModel(data)
    model = AslRestModel(dm, tau=1.8, casl=True, plds=PLDS, repeats=[1])
    fit = SvbFit(dm, model)
    n_it = 400
    fit._setup(model.tpts(), dm.data_flattened, None, 10, 0.05, epochs=n_it, force_num_latent_loss=True)
    f = fit.fused
    state0 = f.state.cpu().numpy().astype(np.float64)
    cfg = om.AslConfig(casl=True, tau=1.8, t1b=1.65)
    spec = H.aslrest_spec(cfg)
    eps_all = []
    for it in range(n_it):
        eps_all.append(f.fill_eps(it).cpu().numpy())
        f.step(1)
    gpu_state = f.state.cpu().numpy()
    ost, _ = eng.fit(spec, torch.as_tensor(state0), torch.zeros(0, dtype=torch.float64),
                     torch.as_tensor(data.T.astype(np.float64)), torch.as_tensor(model.tpts().T.astype(np.float64)),
                     n_it, 6, 0.05, lambda it: torch.as_tensor(eps_all[it], dtype=torch.float64))
    ost = ost.numpy()
    rel_f = np.abs(gpu_state[0] - ost[0]) / np.abs(ost[0])
    rel_d = np.abs(gpu_state[1] - ost[1]) / np.abs(ost[1])
    assert np.median(rel_f) < 1e-4 and np.median(rel_d) < 1e-4
    # a few voxels whose arrival time is not identified by the data (delttiss beyond the last PLD) follow a
    # chaotic Adam trajectory in float32 vs float64; everything that has converged agrees within 1e-3
    assert np.mean(rel_f < 1e-3) > 0.97 and np.mean(rel_d < 1e-3) > 0.97, (rel_f.max(), rel_d.max())


def test_fused_iterations_equal_single_iterations():
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.svbcompat.fit import SvbFit
    rng = np.random.default_rng(3)
    vol, _f, _d = _sim_volume((4, 8, 9), rng, repeats=8, slicedt=0.0452)
    data = vol.reshape(-1, 48)
    # 8 and 16 per launch: the per-iteration cost sums leave the CTA after the iteration loop; 24: more than the kernel
    # defers (kMaxDeferredCosts), reduced inside the loop; state and moments are written back once per launch in all
    states, moments, cost_hists = [], [], []
    n_total = 48
    for fuse in (1, 8, 16, 24):
        dm = DataModel(data)
        model = AslRestModel(dm, tau=1.8, casl=True, plds=PLDS, repeats=[8], slicedt=0.0452, inferart=True)
        fit = SvbFit(dm, model)
        fit._setup(model.tpts(), dm.data_flattened, 6, 10, 0.01, epochs=8, force_num_latent_loss=True)
        f = fit.fused
        assert f.n_batches == 8 and f.B == 6
        # after the posterior initialisation (which reads the data): non-finite samples in the device copy
        f.data[3, 5] = float("nan")         # a voxel whose update is skipped in every iteration that sees row 3
        f.data[7, 9] = float("nan")         # ... and one skipped in the LAST iteration of each fused launch
        for _ in range(n_total // fuse):
            f.step(fuse)
        assert f.step_count == n_total
        states.append(f.state.cpu().numpy())
        moments.append((f.m.cpu().numpy(), f.v.cpu().numpy()))
        costs = f.cost_hist[:n_total].cpu().numpy()
        assert np.isfinite(costs).all() and (costs != 0).all()
        cost_hists.append(costs)
        assert int(f.nan_count.item()) == 12                   # 48 iterations over 8 strided batches: each row six times
    for k in range(1, len(states)):
        np.testing.assert_array_equal(states[0], states[k])
        # the Adam moments stay in shared memory between fused iterations and must come back identical
        np.testing.assert_array_equal(moments[0][0], moments[k][0])
        np.testing.assert_array_equal(moments[0][1], moments[k][1])
        # per-iteration cost sums: the same per-CTA floats, added in double by atomics (order across CTAs is free)
        np.testing.assert_allclose(cost_hists[0], cost_hists[k], rtol=1e-12)
    assert np.isfinite(states[0]).all()


def test_asl_example_real_data_options_on_synthetic_volume(tmp_path):
    """scripts/asl_example.py:24-42 option dict (6 PLD x 8 repeats, slicedt, batch_size 6, lr 0.01) + mask."""
    from svb.main import run
    from svb_models_asl_b200.svbcompat import nifti
    rng = np.random.default_rng(4)
    vol, ftiss, delt = _sim_volume((8, 8, 6), rng, repeats=8, noise=1.0, t1b=1.65, slicedt=0.0452)
    mask = (rng.uniform(size=(8, 8, 6)) < 0.7).astype(np.int16)
    nifti.save(vol, str(tmp_path / "asldata_diff.nii.gz"))
    nifti.save(mask, str(tmp_path / "asldata_mask.nii.gz"))
    options = {"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [8], "slicedt": 0.0452, "learning_rate": 0.01,
               "batch_size": 6, "sample_size": 10, "epochs": 300, "save_mean": True, "save_var": True,
               "force_num_latent_loss": True, "train_load": "trained_data", "display_step": 0,
               "iters_per_launch": 8}
    out = str(tmp_path / "asl_example_out")
    runtime, svb, history = run(str(tmp_path / "asldata_diff.nii.gz"), "aslrest", out,
                                mask=str(tmp_path / "asldata_mask.nii.gz"), **options)
    mf = nifti.load(os.path.join(out, "mean_ftiss.nii.gz")).data
    md = nifti.load(os.path.join(out, "mean_delttiss.nii.gz")).data
    assert mf.shape == (8, 8, 6)
    assert (mf[mask == 0] == 0).all() and (mf[mask > 0] != 0).all()
    sel = mask > 0
    assert np.median(np.abs(mf[sel] - ftiss[sel]) / ftiss[sel]) < 0.08
    assert np.median(np.abs(md[sel] - delt[sel])) < 0.15
    assert svb.fused.step_count == 300 * 8


def test_spatial_prior_fit_smooths_and_learns_ak(tmp_path):
    """param_overrides {"ftiss": {"prior_type": "M"}}: the MRF prior runs end to end on one GPU."""
    from svb.main import run
    from svb_models_asl_b200.svbcompat import nifti
    rng = np.random.default_rng(5)
    shape = (8, 8, 8)
    vol, ftiss, delt = _sim_volume(shape, rng, noise=3.0, t1b=1.65)
    nifti.save(vol, str(tmp_path / "sig.nii.gz"))
    base = {"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "learning_rate": 0.05, "sample_size": 10,
            "epochs": 400, "save_mean": True, "force_num_latent_loss": True, "display_step": 0}
    _rt, svb_n, _h = run(str(tmp_path / "sig.nii.gz"), "aslrest", str(tmp_path / "n"), **base)
    _rt, svb_m, hist = run(str(tmp_path / "sig.nii.gz"), "aslrest", str(tmp_path / "m"),
                           param_overrides={"ftiss": {"prior_type": "M"}}, **base)
    assert svb_m.fused.mrf == [0]
    lak = float(svb_m.fused.log_ak[0])
    assert math.isfinite(lak) and lak != pytest.approx(math.log(1e-5))
    fn = nifti.load(str(tmp_path / "n" / "mean_ftiss.nii.gz")).data
    fm = nifti.load(str(tmp_path / "m" / "mean_ftiss.nii.gz")).data
    assert np.isfinite(fm).all() and hist["nan_skips"] == 0

    def rough(v):
        return sum(np.abs(np.diff(v, axis=a)).mean() for a in range(3))
    assert rough(fm) <= rough(fn) * 1.001          # never rougher than the voxel-wise fit


def test_aslnn_and_disp_fits_run_through_the_plugin_api(tmp_path):
    """aslnn (train_load weights in the reference .npy layout, asl_example_sim_nn.py:23-41) and aslrest_disp
    drive the same engine through run(); outputs are finite and the cost decreases."""
    from svb.main import run
    from svb_models_asl_b200.svbcompat import nifti
    rng = np.random.default_rng(6)
    vol, ftiss, delt = _sim_volume((6, 6, 6), rng, noise=0.3, t1b=1.65)
    nifti.save(vol, str(tmp_path / "sig.nii.gz"))
    # weights: a quick fit of the MLP to the analytic curve (scripts/retrain_model.py does this properly)
    wdir = str(tmp_path / "trained_data")
    os.makedirs(wdir)
    g = np.random.default_rng(0)
    for i, (a, b) in enumerate([(2, 10), (10, 10), (10, 1)]):
        np.save(os.path.join(wdir, "weights%i.npy" % i), g.normal(0, 0.5, (a, b)).astype(np.float32))
        np.save(os.path.join(wdir, "biases%i.npy" % i), g.normal(0, 0.1, (1, b)).astype(np.float32))
    base = {"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "learning_rate": 0.05, "sample_size": 10,
            "epochs": 200, "save_mean": True, "save_model_fit": True, "force_num_latent_loss": True, "display_step": 0}
    _rt, svb_nn, h_nn = run(str(tmp_path / "sig.nii.gz"), "aslnn", str(tmp_path / "nn"), train_load=wdir, **base)
    assert [p.name for p in svb_nn.params] == ["ftiss", "delttiss", "noise"]
    assert np.isfinite(h_nn["mean_cost"]).all() and h_nn["mean_cost"][-1] < h_nn["mean_cost"][0]
    mf = nifti.load(str(tmp_path / "nn" / "mean_ftiss.nii.gz")).data
    assert np.isfinite(mf).all() and (mf > 0).all()                      # LogNormal: exp(mean)
    _rt, svb_d, h_d = run(str(tmp_path / "sig.nii.gz"), "aslrest_disp", str(tmp_path / "disp"), inferart=True, **base)
    assert [p.name for p in svb_d.params] == ["ftiss", "delttiss", "fblood", "deltblood", "s", "sp", "noise"]
    assert np.isfinite(h_d["mean_cost"]).all() and h_d["mean_cost"][-1] < h_d["mean_cost"][0]
    fit = nifti.load(str(tmp_path / "disp" / "modelfit.nii.gz")).data
    assert fit.shape == (6, 6, 6, 6) and np.isfinite(fit).all()
    assert np.isfinite(nifti.load(str(tmp_path / "disp" / "mean_delttiss.nii.gz")).data).all()


def test_disp_fit_recovers_parameters_of_dispersed_data(tmp_path):
    """Data generated BY the dispersion model (fixed Fabber-default s, sp) are recovered by fitting it."""
    from svb import DataModel
    from svb.main import run
    from svb_models_asl import AslRestDisp
    from svb_models_asl_b200.svbcompat import nifti
    rng = np.random.default_rng(8)
    n = 6 * 6 * 6
    ftiss, delt = rng.uniform(2.0, 20.0, n), rng.uniform(0.6, 2.2, n)
    gen = AslRestDisp(DataModel(np.zeros((1, 6), dtype=np.float32)), tau=1.8, casl=True, plds=PLDS, repeats=1,
                      infer_disp_params=False)
    params = np.stack([ftiss, delt]).astype(np.float32).reshape(2, n, 1, 1)
    t = np.asarray(gen.tis, dtype=np.float32).reshape(1, 1, -1)
    sig = gen.ievaluate(params, t)[:, 0, :] + rng.normal(0, 0.1, (n, 6))
    nifti.save(sig.reshape(6, 6, 6, 6).astype(np.float32), str(tmp_path / "sig.nii.gz"))
    _rt, svb, hist = run(str(tmp_path / "sig.nii.gz"), "aslrest_disp", str(tmp_path / "out"), tau=1.8, casl=True,
                         plds=PLDS, repeats=[1], infer_disp_params=False, learning_rate=0.05, sample_size=10,
                         epochs=600, save_mean=True, force_num_latent_loss=True, display_step=0)
    assert [p.name for p in svb.params] == ["ftiss", "delttiss", "noise"]
    mf = nifti.load(str(tmp_path / "out" / "mean_ftiss.nii.gz")).data.ravel()
    md = nifti.load(str(tmp_path / "out" / "mean_delttiss.nii.gz")).data.ravel()
    assert np.median(np.abs(mf - ftiss) / ftiss) < 0.05
    assert np.median(np.abs(md - delt)) < 0.06


def test_host_fed_steps_equal_device_resident_steps():
    """svbasl_step_host (batches from pinned host memory, full or low-rank time points) follows exactly the same
    trajectory as device-resident svbasl_step."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.ops import HostFeeder
    from svb_models_asl_b200.svbcompat.fit import SvbFit
    rng = np.random.default_rng(12)
    vol, _f, _d = _sim_volume((5, 7, 6), rng, noise=0.5, t1b=1.65, slicedt=0.0452)
    data = vol.reshape(-1, 6)
    states = []
    for mode in ("device", "host_full", "host_lowrank"):
        dm = DataModel(vol)
        model = AslRestModel(dm, tau=1.8, casl=True, plds=PLDS, repeats=[1], slicedt=0.0452, inferart=True)
        fit = SvbFit(dm, model)
        fit._setup(model.tpts(), dm.data_flattened, None, 10, 0.05, epochs=20, force_num_latent_loss=True)
        f = fit.fused
        if mode == "device":
            for _ in range(7):
                f.step(1)
        else:
            feeder = HostFeeder(f)
            h_data = torch.from_numpy(np.ascontiguousarray(dm.data_flattened.T)).pin_memory()
            if mode == "host_full":
                h_t = torch.from_numpy(np.ascontiguousarray(model.tpts().T)).pin_memory()
                for _ in range(7):
                    feeder.step(h_data, host_tpts=h_t)
            else:
                ti, zoff = model.tpts_lowrank()
                h_ti = torch.from_numpy(ti).pin_memory()
                zd = torch.as_tensor(zoff, device=f.dev)
                for _ in range(7):
                    feeder.step(h_data, host_ti=h_ti, zoff_dev=zd)
            cost = feeder.sync()
            assert np.isfinite(cost)
            feeder.close()
        assert f.step_count == 7
        states.append(f.state.cpu().numpy())
    np.testing.assert_array_equal(states[0], states[1])
    # ti + z*slicedt is formed in float32 on the device instead of float64-then-rounded on the host: 1 ulp in t
    np.testing.assert_allclose(states[2], states[0], rtol=2e-4, atol=2e-5)


def test_device_initialisers_match_reference_goldens(golden):
    """svbasl_init_stats (the per-voxel reductions behind _init_flow / _init_fblood / att_init="max",
    aslrest.py:461-520) on the real-data sample of the golden fixture, through the plugin's own callbacks: the
    device-reduced path (ops.InitData.device_stats, what SvbFit._setup hands to post_init) reproduces the values
    the unmodified reference computed with numpy."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.ops import InitData, device_array, device_init_stats
    g = golden("aslrest_real")["real"]
    data, tpts = g["data_sel"], g["tpts_sel"]                        # [343, 48] float32
    dev = torch.device("cuda:0")
    stats = device_init_stats(device_array(data.T.copy(), dev), device_array(tpts.T.copy(), dev))
    stats = {k: v.cpu().numpy() for k, v in stats.items()}
    np.testing.assert_allclose(stats["mean_t"], data.astype(np.float64).mean(1), rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(stats["max_t"], data.max(1))
    np.testing.assert_allclose(stats["var_t"], data.astype(np.float64).var(1), rtol=1e-5)
    np.testing.assert_array_equal(stats["t_at_max"], tpts[np.arange(len(data)), data.argmax(1)])
    dm = DataModel(data)
    init = InitData(data, stats)
    for att_init, key_mean, key_var in (("", None, None), ("max", "init_delt_max_sel", "init_delt_max_var_sel")):
        model = AslRestModel(dm, tau=1.8, casl=True, plds=PLDS, repeats=[8], slicedt=0.0452, inferart=True,
                             att_init=att_init)
        par = {p.name: p for p in model.params}
        np.testing.assert_allclose(par["ftiss"].post_init(par["ftiss"], tpts, init)[0], g["init_ftiss_sel"], rtol=2e-6)
        np.testing.assert_array_equal(par["fblood"].post_init(par["fblood"], tpts, init)[0], g["init_fblood_sel"])
        mean, var = par["delttiss"].post_init(par["delttiss"], tpts, init)
        if key_mean:
            np.testing.assert_allclose(mean, g[key_mean], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(var, g[key_var], rtol=1e-6)
        else:
            np.testing.assert_allclose(mean, float(g["init_delt"]), rtol=1e-6)
            np.testing.assert_allclose(var, float(g["init_delt_var"]), rtol=1e-6)
        # the host path (plain ndarray, as the reference passes it) gives the same initial posterior
        np.testing.assert_allclose(par["ftiss"].post_init(par["ftiss"], tpts, data)[0], g["init_ftiss_sel"], rtol=2e-6)


def test_nn_trainer_reduces_loss_and_round_trips_weights(tmp_path):
    """AslNNModel's SGD trainer (aslnn.py:155-170, 262-299) on AslRestModel-simulated curves and the reference's
    weights%i.npy / biases%i.npy layout (aslnn.py:326-340): a short run fits better than the initial network,
    saves six files with the reference shapes and a fresh model loading them evaluates identically."""
    from svb import DataModel
    from svb_models_asl import AslNNModel
    dm = DataModel(np.zeros((1, 6), dtype=np.float32))
    opts = dict(tau=1.8, t1b=1.6, casl=True, repeats=1, t1=1.3, tis=[1.8 + p for p in PLDS])
    model = AslNNModel(dm, train_examples=4000, train_steps=60, train_lr=0.05, train_batch_size=200,
                       train_save=str(tmp_path / "w"), **opts)
    x_train, x_test, y_train, y_test = model._get_training_data(4000, seed=3)
    model._train_nn(x_train, y_train, 1, 0.05, 200)
    mse0 = float(np.mean((model._ievaluate_nn(x_test) - y_test) ** 2))
    model._train_nn(x_train, y_train, 60, 0.05, 200)
    mse1 = float(np.mean((model._ievaluate_nn(x_test) - y_test) ** 2))
    assert mse1 < 0.5 * mse0, (mse0, mse1)
    model._save_nn(str(tmp_path / "w"))
    shapes = [np.load(str(tmp_path / "w" / n)).shape for n in
              ("weights0.npy", "biases0.npy", "weights1.npy", "biases1.npy", "weights2.npy", "biases2.npy")]
    assert shapes == [(2, 10), (1, 10), (10, 10), (1, 10), (10, 1), (1, 1)]
    again = AslNNModel(dm, train_load=str(tmp_path / "w"), **opts)
    np.testing.assert_array_equal(again._ievaluate_nn(x_test), model._ievaluate_nn(x_test))


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_spatial_iteration_follows_the_oracle(use_graph, record_error):
    """One launch per iteration with an MRF prior (svbasl_step_spatial: neighbour samples read from the buffer the
    previous launch wrote, next-iteration samples written after the Adam update, log-ak gradient reduced and stepped
    by the last CTA, iteration counter on the device; eager and as a CUDA-graph replay) against the oracle's fit with
    a trainable log ak (autograd through the Laplacian term) on identical draws."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.svbcompat.fit import SvbFit
    rng = np.random.default_rng(17)
    shape = (6, 5, 4)
    vol, _f, _d = _sim_volume(shape, rng, noise=1.0, t1b=1.65)
    dm = DataModel(vol)
    over = {"ftiss": {"prior_type": "M"}}
    model = AslRestModel(dm, tau=1.8, casl=True, plds=PLDS, repeats=[1], param_overrides=over)
    fit = SvbFit(dm, model)
    fit._setup(model.tpts(), dm.data_flattened, None, 10, 0.05, epochs=16, force_num_latent_loss=True, use_graph=False,
               ak=0.3, seed=5, param_overrides=over)
    f = fit.fused
    W, n_it = dm.n_nodes, 5
    cfg = om.AslConfig(casl=True, tau=1.8, t1b=1.65)
    spec = H.aslrest_spec(cfg, mrf=(0,))
    e = f.engine_desc()
    assert [e.prior_type[i] for i in range(3)] == [2, 0, 0]
    np.testing.assert_allclose([e.prior_mean[i] for i in range(3)], [float(np.mean(x)) for x in spec.prior_mean], rtol=1e-6)
    np.testing.assert_allclose([e.prior_var[i] for i in range(3)], spec.prior_var, rtol=1e-6)
    prob = H.synth_problem(cfg, spec, W, rng)                      # only its random posterior state is used
    state0 = prob["state"].astype(np.float32)
    f.state.copy_(torch.as_tensor(state0, device=f.dev))
    f.sp_valid = False
    eps_all = [f.fill_eps(it).cpu().numpy() for it in range(n_it)]
    data = dm.data_flattened.T.astype(np.float64)
    tp = np.broadcast_to(model.tpts(), dm.data_flattened.shape).T.astype(np.float64)
    ost, ohy = eng.fit(spec, torch.as_tensor(state0.astype(np.float64)), torch.tensor([math.log(0.3)], dtype=torch.float64),
                       torch.as_tensor(data), torch.as_tensor(tp), n_it, 6, 0.05,
                       lambda it: torch.as_tensor(eps_all[it], dtype=torch.float64),
                       neighbours=torch.as_tensor(dm.neighbour_table().astype(np.int64)))
    if use_graph:
        f.enable_graph()
    for _ in range(n_it):
        f.step()
    f.check_peers()
    assert int(f.step_dev.item()) == n_it and f.step_count == n_it
    st = f.state.cpu().numpy()
    stats = H.trajectory_error(st, ost.numpy())
    lak_err = abs(float(f.log_ak[0]) - float(ohy[0])) / abs(float(ohy[0]))
    record_error("fused_spatial_vs_oracle/%s" % ("graph" if use_graph else "eager"), log_ak_rel=float(lak_err), **stats)
    assert np.abs(ost.numpy() - state0).max() > 0.05 and abs(float(ohy[0]) - math.log(0.3)) > 0.05
    assert stats["q50"] <= 2e-6 and stats["q90"] <= 1e-4 and stats["q99"] <= 1e-3, stats   # the bulk
    assert stats["max_abs"] <= 2 * 0.05 * n_it, stats                          # stragglers: Adam's step envelope
    assert lak_err <= 1e-4, lak_err
    costs = f.cost_hist[:n_it].cpu().numpy()
    assert np.isfinite(costs).all() and (costs != 0).all()


def test_device_resident_setup_equals_host_setup():
    """SvbFit.setup_from_device (data already on the GPU, low-rank time points, neighbour rows of the local range only
    - what bench.py's 10 M-voxel volume goes through) sets up the same fit as the host-array path of train(): same
    initial posterior, same state and log ak after 12 iterations with a spatial prior."""
    from svb import DataModel
    from svb_models_asl import AslRestModel
    from svb_models_asl_b200.sharding import ShardPlan
    from svb_models_asl_b200.svbcompat.fit import SvbFit
    rng = np.random.default_rng(23)
    shape = (7, 6, 5)
    vol, _f, _d = _sim_volume(shape, rng, repeats=8, noise=1.0, t1b=1.65, slicedt=0.0452)
    over = {"ftiss": {"prior_type": "M"}}
    opts = dict(tau=1.8, casl=True, plds=PLDS, repeats=[8], slicedt=0.0452, param_overrides=over)
    results = []
    for route in ("host", "device"):
        dm = DataModel(vol) if route == "host" else DataModel.header(shape, 48)
        model = AslRestModel(dm, **opts)
        fit = SvbFit(dm, model)
        if route == "host":
            fit._setup(model.tpts(), dm.data_flattened, 6, 10, 0.01, epochs=4, force_num_latent_loss=True, param_overrides=over)
            f = fit.fused
        else:
            plan = ShardPlan(dm.n_nodes, 0, 1, dm.neighbour_table)
            dev = torch.device("cuda:0")
            data = torch.as_tensor(vol.reshape(-1, 48).T.copy(), device=dev)
            ti, zoff = model.tpts_lowrank()
            f = fit.setup_from_device(data, plan, ti, torch.as_tensor(zoff, device=dev), 6, 10, 0.01, 40, param_overrides=over)
        init = f.state.cpu().numpy().copy()
        for _ in range(12):
            f.step()
        f.check_peers()
        results.append((init, f.state.cpu().numpy(), float(f.log_ak[0])))
        f.release()
    np.testing.assert_allclose(results[1][0], results[0][0], rtol=2e-6, atol=1e-6)      # initial posterior
    np.testing.assert_allclose(results[1][1], results[0][1], rtol=2e-4, atol=2e-5)      # full [T, W] time points vs ti + zoff
    assert results[1][2] == pytest.approx(results[0][2], rel=1e-5)
