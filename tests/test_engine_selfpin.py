"""
CPU tier: the engine oracle reproduces its own frozen outputs (tests/golden/engine_selfpin.npz, written by
tests/golden/make_engine_selfpin.py).  This is a drift guard for the checker, NOT a pin to the reference - the
engine half of the oracle stays "parity unpinned" (DESIGN.md section 6).
"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_engine_selfpin", os.path.join(HERE, "golden", "make_engine_selfpin.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


@pytest.mark.parametrize("case", sorted(gen.CASES))
def test_engine_oracle_reproduces_its_frozen_outputs(case):
    frozen = np.load(os.path.join(HERE, "golden", "engine_selfpin.npz"))
    now = gen.compute(case)
    for key, value in now.items():
        want = frozen["%s/%s" % (case, key)]
        assert value.shape == want.shape
        if value.size == 0:
            continue
        np.testing.assert_allclose(value, want, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(want).max()), err_msg=key)
        assert np.isfinite(value).all()
