"""
aslnn on the tensor cores (tcgen05.mma kind::tf32 + TMEM + TMA bulk copy, csrc/nn_tc.cu) against
(a) the golden vectors of the reference source, (b) the FP32-pipe kernel, (c) an fp64 numpy GEMM for the hidden
pre-activations - tolerance 1e-5 relative (BASELINE.json forward tolerance); ragged / tiny / broadcast shapes.
"""
import numpy as np
import pytest
import torch

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
