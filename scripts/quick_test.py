#!/usr/bin/env python
"""
Known-answer check of the kinetic curve - the reference's scripts/quick_test.py configuration (:10-24: buxton, tau 1.8,
t1b 1.6, t1 1.3, CASL, TIs 2.05..3.3; ftiss = 1, 5, 10 at delttiss = 1.3, :27-33,42-49).  The reference prints Fabber's
`model_evaluate` next to the NN surrogate's `ievaluate` for eyeballing; Fabber is not available here, so the analytic
side is AslRestModel.ievaluate (the same Buxton curve, aslrest.py:342-391) AND the stored known answer (SURVEY.md
Appendix D1: float64 restatement of the formula), and both comparisons are asserted instead of printed only.

    python scripts/quick_test.py        (needs a GPU: the curves come from the CUDA evaluate kernel)
"""
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svb import DataModel  # noqa: E402
from svb_models_asl import AslNNModel, AslRestModel  # noqa: E402

tis = [2.05, 2.3, 2.55, 2.8, 3.05, 3.3]
options = {                      # scripts/quick_test.py:10-24 (Fabber-style keys are ignored by the svb models)
    "model" : "buxton",
    'lambda': 0.9,
    'tau' : 1.8,
    'ti' : tis,
    't1b': 1.6,
    "prior-noise-stddev" : 1,
    'casl': True,
    'repeats': 1,
    't1': 1.3
}
# ftiss = 1, delttiss = 1.3
KNOWN_F1 = np.array([0.503872642, 0.616141242, 0.708511842, 0.784511077, 0.847040537, 0.734146299])


def main():
    logging.getLogger().setLevel(logging.INFO)
    sig = np.zeros((1, 6), dtype=np.float32)
    data_model = DataModel(sig)
    tpts = np.zeros((1, 6), dtype=np.float32)
    tpts[..., :] = tis
    # 2 params, 3 voxels, 1 samples                                     (quick_test.py:42-47)
    params = np.zeros((2, 3, 1), dtype=np.float32)
    params[0, 0, :] = 1.0
    params[0, 1, :] = 5.0
    params[0, 2, :] = 10.0
    params[1, :, :] = 1.3
    analytic = AslRestModel(data_model, tis=tis, **options).ievaluate(params, tpts)
    print("analytic (aslrest kernel):")
    print(analytic)
    for row, ftiss in enumerate((1.0, 5.0, 10.0)):
        np.testing.assert_allclose(analytic[row].ravel(), ftiss * KNOWN_F1, rtol=1e-5)
    wdir = "trained_data" if os.path.isdir("trained_data") else os.path.join(ROOT, "trained_data")
    model = AslNNModel(data_model, tis=tis, train_load=wdir, **options)
    modelsig = model.ievaluate(params, tpts)
    print("NN surrogate (aslnn kernel):")
    print(modelsig)
    err = np.abs(modelsig - analytic).max() / np.abs(analytic).max()
    print("max |NN - analytic| / max |analytic| = %.4f" % err)
    assert err < 0.03, "surrogate does not follow the analytic curve"
    print("quick_test OK")


if __name__ == "__main__":
    main()
