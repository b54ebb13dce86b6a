"""The C-ABI library loads and exports every symbol include/svbasl.h declares (no compute calls: CPU tier)."""
import ctypes
import os
import re

import pytest

from svb_models_asl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "svbasl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svbasl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from svb_models_asl_b200.build import build
        build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.svbasl_abi_version.restype = ctypes.c_int
    assert lib.svbasl_abi_version() == 2


def test_struct_layouts_match_the_header():
    """Field order/size drift between svbasl.h and the ctypes mirror would corrupt every call."""
    import subprocess
    import tempfile
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "svbasl.h"
int main(void) {
  printf("%zu %zu %zu ", sizeof(svbasl_model), sizeof(svbasl_engine), sizeof(svbasl_adam));
  printf("%zu %zu %zu %zu ", offsetof(svbasl_model, pvgm), offsetof(svbasl_model, nn_weights),
         offsetof(svbasl_engine, state), offsetof(svbasl_engine, ak_grad));
  printf("%zu %zu ", offsetof(svbasl_engine, prior_var), offsetof(svbasl_adam, step0));
  printf("%zu %zu %zu %zu %zu\n", sizeof(svbasl_hyper), offsetof(svbasl_hyper, mailboxes), offsetof(svbasl_hyper, status),
         offsetof(svbasl_engine, peer_hi_count), offsetof(svbasl_engine, cost_sum_scalar));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "l.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "l")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    M, E, A, Hy = _lib.Model, _lib.Engine, _lib.Adam, _lib.Hyper
    want = [ctypes.sizeof(M), ctypes.sizeof(E), ctypes.sizeof(A), M.pvgm.offset, M.nn_weights.offset, E.state.offset,
            E.ak_grad.offset, E.prior_var.offset, A.step0.offset, ctypes.sizeof(Hy), Hy.mailboxes.offset,
            Hy.status.offset, E.peer_hi_count.offset, E.cost_sum_scalar.offset]
    assert got == want


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from svb import DataModel
    from svb_models_asl import AslRestModel
    dm = DataModel(np.zeros((3, 6), dtype=np.float32))
    model = AslRestModel(dm, tis=[2.05, 2.3, 2.55, 2.8, 3.05, 3.3], casl=True)
    with pytest.raises(_lib.SvbAslError):
        model.evaluate([np.ones((3, 1, 1)), np.ones((3, 1, 1))], np.ones((3, 1, 6)))
