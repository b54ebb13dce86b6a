"""
CPU tier: the running incomplete-gamma evaluator of the dispersion model (csrc/model_disp.h: gamma_run_eval - fixed
14-term series below x = 2, Gauss-Legendre increments above, igammac_d for very wide steps) against scipy's
gammaincc in float64, on the argument sequences the convolution grid produces: x_k = s (k h + r).
The reference gets these values from tf.math.igammac (aslrest_disp.py:91-108).
"""
import ctypes as C

import numpy as np
import pytest
from scipy import special

from tests.hostsim.build_hostsim import build as build_hostsim


@pytest.fixture(scope="module")
def lib():
    lib = C.CDLL(build_hostsim())
    lib.hostsim_gamma_run.argtypes = [C.c_float] + [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    lib.hostsim_gamma_run.restype = None
    return lib


def _run(lib, a, xs):
    xs = np.ascontiguousarray(xs, dtype=np.float32)
    outs = [np.zeros_like(xs) for _ in range(6)]
    lib.hostsim_gamma_run(a, xs.ctypes.data, len(xs), *[o.ctypes.data for o in outs])
    return outs


def _exact(a, xs):
    xs = xs.astype(np.float64)
    q = special.gammaincc(a, xs)
    da = 1e-5
    dqa = (special.gammaincc(a + da, xs) - special.gammaincc(a - da, xs)) / (2 * da)
    with np.errstate(divide="ignore", invalid="ignore"):
        dqx = -np.exp((a - 1) * np.log(xs) - xs - special.gammaln(a))
    return q, dqa, np.where(xs > 0, dqx, 0.0)


@pytest.mark.parametrize("a", [1.05, 1.3, 2.0, 3.7, 6.5, 11.0])
@pytest.mark.parametrize("s", [1.5, 4.0, 7.4, 25.0, 60.0, 120.0])
def test_running_evaluator_matches_scipy_on_grid_sequences(lib, a, s):
    """Grid of 51 points, h = 0.1, arbitrary sub-grid offset; s*h from 0.15 (many series points) over 2.5 (several
    quadrature pieces per step) to 12 (wider than 8: continued-fraction fallback)."""
    rng = np.random.default_rng(int(a * 100 + s))
    xs = (s * (np.arange(51) * 0.1 + rng.uniform(0, 0.1))).astype(np.float32)
    q, dqa, dqx, q0, dqa0, dqx0 = _run(lib, a, xs)
    eq, edqa, edqx = _exact(a, xs)
    # absolute accuracy relative to the function's range (Q in [0,1]; |dQ/da| <~ 0.5; density <~ 1)
    assert np.abs(q - eq).max() < 2e-6, np.abs(q - eq).max()
    assert np.abs(dqa - edqa).max() < 5e-6, np.abs(dqa - edqa).max()
    assert np.abs(dqx - edqx).max() < 2e-6 * max(1.0, np.abs(edqx).max())
    # and no worse than twice the from-scratch evaluation it replaces (+ float32 accumulation of 50 increments)
    assert np.abs(q - eq).max() <= 2 * np.abs(q0 - eq).max() + 1.5e-6
    assert np.abs(dqa - edqa).max() <= 2 * np.abs(dqa0 - edqa).max() + 3e-6


def test_restart_when_the_argument_goes_down_and_at_zero(lib):
    a = 2.4
    xs = np.asarray([0.0, 0.3, 5.0, 5.5, 1.0, 9.0, 9.0, 30.0, 3.0], dtype=np.float32)   # not monotone, repeats, x = 0
    q, dqa, dqx, *_ = _run(lib, a, xs)
    eq, edqa, edqx = _exact(a, xs)
    assert q[0] == 1.0 and dqa[0] == 0.0 and dqx[0] == 0.0
    np.testing.assert_allclose(q, eq, atol=2e-6)
    np.testing.assert_allclose(dqa, edqa, atol=5e-6)
    np.testing.assert_allclose(dqx, edqx, atol=2e-6)
