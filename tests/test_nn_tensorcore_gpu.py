"""
aslnn on the tensor cores (tcgen05.mma kind::tf32 + TMEM + TMA bulk copy, csrc/nn_tc.cu) against
(a) the golden vectors of the reference source, (b) the FP32-pipe kernel, (c) an fp64 numpy GEMM for the hidden
pre-activations - tolerance 1e-5 relative (BASELINE.json forward tolerance); ragged / tiny / broadcast shapes.
"""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _model(ws, bs, tmp_path):
    from svb import DataModel
    from svb_models_asl import AslNNModel
    for i, (w, b) in enumerate(zip(ws, bs)):
        np.save(tmp_path / ("weights%i.npy" % i), w)
        np.save(tmp_path / ("biases%i.npy" % i), b)
    dm = DataModel(np.zeros((1, 6), dtype=np.float32))
    return AslNNModel(dm, tis=[2.05, 2.3, 2.55, 2.8, 3.05, 3.3], casl=True, train_load=str(tmp_path))


def test_tensor_core_mlp_matches_reference_golden(golden, tmp_path):
    n = golden("aslnn_eval")["nn"]
    ws, bs = [n["w%i" % i] for i in range(3)], [n["b%i" % i] for i in range(3)]
    model = _model(ws, bs, tmp_path)
    model.use_tensor_cores = True
    out = model.evaluate(list(n["params"]), n["t"]).cpu().numpy()          # tensor-core path
    ref = n["out64"]
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()
    model.use_tensor_cores = False
    simt = model.evaluate(list(n["params"]), n["t"]).cpu().numpy()
    assert np.abs(out - simt).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("W,S,B", [(1, 1, 6), (5, 3, 6), (1000, 10, 6), (257, 4, 48), (40000, 10, 6)])
def test_tensor_core_hidden_layer_is_float32_accurate(tmp_path, W, S, B):
    from svb_models_asl_b200.ops import nn_evaluate_tc
    rng = np.random.default_rng(W + B)
    ws = [rng.normal(0, 0.8, s).astype(np.float32) for s in [(2, 10), (10, 10), (10, 1)]]
    bs = [rng.normal(0, 0.3, (1, s)).astype(np.float32) for s in (10, 10, 1)]
    model = _model(ws, bs, tmp_path)
    f = rng.uniform(0.5, 20, (W, S, 1)).astype(np.float32)
    d = rng.uniform(0.1, 3.0, (W, S, 1)).astype(np.float32)
    t = (rng.uniform(1.0, 5.0, (W, 1, B))).astype(np.float32)
    out, hidden = nn_evaluate_tc(model, [f, d], t, want_hidden=True)
    out, hidden = out.cpu().numpy(), hidden.cpu().numpy()
    x = np.stack(np.broadcast_arrays(t.astype(np.float64), d.astype(np.float64)), -1)      # [W,S,B,2]
    h1 = np.tanh(x @ ws[0].astype(np.float64) + bs[0].astype(np.float64))
    z2 = h1 @ ws[1].astype(np.float64) + bs[1].astype(np.float64)
    ref = f.astype(np.float64) * (np.tanh(z2) @ ws[2].astype(np.float64) + bs[2].astype(np.float64))[..., 0]
    assert np.abs(hidden - z2).max() <= 2e-6 * np.abs(z2).max()            # 3-term tf32 split: ~2^-20
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


def test_tensor_core_sass_is_blackwell_native():
    """The library holds tcgen05 MMA / TMEM load / TMA bulk-copy instructions (B200_PROFILING.md mnemonics)."""
    import os
    import shutil
    import subprocess
    from svb_models_asl_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def _nn_problem(golden, W, rng):
    import torch
    from oracle import asl_models as om
    n = golden("aslnn_eval")["nn"]
    ws, bs = [n["w%i" % i] for i in range(3)], [n["b%i" % i] for i in range(3)]
    spec = H.nn_spec(ws, bs, latent="numeric")
    tpts = np.repeat(np.asarray(H.TIS, dtype=np.float32)[:, None], W, 1)
    state = np.stack([rng.normal(1.5, 0.5, W), rng.normal(1.2, 0.6, W) * rng.choice([-1, 1], W), rng.normal(0, 0.3, W)]
                     + [rng.normal(-2, 0.5, W) for _ in range(3)] + [rng.normal(0, 0.05, W) for _ in range(3)])
    state = state.astype(np.float32).astype(np.float64)
    f = torch.as_tensor(np.exp(state[0])).reshape(W, 1, 1)
    d = torch.as_tensor(np.abs(state[1])).reshape(W, 1, 1)
    clean = om.evaluate_nn(ws, bs, f, d, torch.as_tensor(tpts.astype(np.float64)).T.unsqueeze(1))[:, 0, :].T.numpy()
    prob = {"state": state, "tpts": tpts, "data": (clean + rng.normal(0, 0.5, clean.shape)).astype(np.float32)}
    return spec, prob


@pytest.mark.parametrize("B", [6, 5, 1])
def test_fused_step_with_tensor_core_products_matches_oracle(golden, record_error, B):
    """The fused ELBO + gradient of aslnn with the two 10x10 products per row on tcgen05 (csrc/model_nn_tc.cuh;
    SVBASL_F_NN_TC): cost and gradient against the oracle (autograd through the MLP, aslnn.py:238-260) at the north-star
    tolerances, on 300 voxels = two full 128-row tiles and a ragged one.  B = 6 is the register-resident instantiation
    (three rounds of two time points), B = 5 and 1 the any-batch-size one with a single-row last round."""
    from tests.test_kernel_parity import _check_grads, _record_grad_errors
    be = H.Backend("cuda")
    rng = np.random.default_rng(44)
    W = 300
    spec, prob = _nn_problem(golden, W, rng)
    if B != 6:
        spec = H.nn_spec(spec.cfg["weights"], spec.cfg["biases"], latent="numeric", t_full=B)
        prob = dict(prob, tpts=prob["tpts"][:B], data=prob["data"][:B])
    eps = rng.normal(size=(3, spec.n_samples, W)).astype(np.float32)
    ocost, ograd, _gh, _ = H.oracle_cost_grad(spec, prob, eps)
    m = be.model_desc(dict(spec.cfg, tensor_cores=True))
    e, _b = be.engine_desc(spec, prob["state"], prob["data"], prob["tpts"], eps)
    cost, grad, _ = be.elbo_grad(m, e, spec.n_state)
    _record_grad_errors(record_error, "nn_tc_fused/B%d/cuda" % B, cost, grad, ocost, ograd)
    _check_grads(cost, grad, ocost, ograd)


def test_production_step_tensor_cores_equals_fp32_pipe(golden):
    """svbasl_step (lean flavour, in-kernel draws, 3 fused iterations per launch) with and without SVBASL_F_NN_TC:
    same posterior after 6 iterations up to float32 rounding of the two product forms."""
    be = H.Backend("cuda")
    rng = np.random.default_rng(45)
    W = 1000
    spec, prob = _nn_problem(golden, W, rng)
    finals = []
    for tc in (False, True):
        m = be.model_desc(dict(spec.cfg, tensor_cores=tc))
        e, bufs = be.engine_desc(spec, prob["state"], prob["data"], prob["tpts"], None, seed=9)
        ad, _ab = be.adam_desc(spec.n_state, W, 0.05, 8, n_iters=3)
        for launch in range(2):
            ad.step0 = 3 * launch
            csum, nanc = be.step(m, e, ad)
            assert nanc == 0 and np.isfinite(csum).all()
        finals.append(be.get(bufs["state"]))
    assert np.abs(finals[0] - prob["state"]).max() > 0.05
    stats = H.trajectory_error(finals[1], finals[0])
    assert stats["q50"] <= 2e-6 and stats["q99"] <= 1e-4 and stats["max_abs"] <= 0.05, stats
