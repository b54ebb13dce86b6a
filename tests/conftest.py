import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        blob = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        tree = {}
        for key in blob.files:
            case, _, field = key.partition("/")
            tree.setdefault(case, {})[field] = blob[key]
        return tree
    return load


_ERRORS = {}


@pytest.fixture
def record_error():
    """Parity tests report the error they measured (not only pass/fail): record_error(name, key=value, ...).
    Written at session end to gpurun_out/parity_errors_<tier>.json; the GPU tier's file is copied to profiles/."""
    def rec(name, **values):
        _ERRORS.setdefault(name, {}).update(values)
    return rec


def pytest_sessionfinish(session, exitstatus):
    if not _ERRORS:
        return
    import json
    try:
        import torch
        tier = "gpu" if torch.cuda.is_available() else "cpu"
    except Exception:  # noqa: BLE001
        tier = "cpu"
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_errors_%s.json" % tier), "w") as f:
        json.dump(_ERRORS, f, indent=1, sort_keys=True)
