#!/usr/bin/env python
"""
Synthetic multi-PLD pCASL test data - the counterpart of the reference's scripts/gen_test_data.py
(TIs :15, ftiss~U(1,20) / delttiss~U(0.6,2.5) :40-41, generator t1b=1.6 :28, NOISE_SD :16), scalable to the
benchmark volumes.  The reference script cannot run as shipped (undefined `tis`/`options`).

    python scripts/gen_test_data.py [--side 10] [--noise 0.0] [--seed 20260101] [--outdir .]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svb import DataModel                                    # noqa: E402
from svb_models_asl import AslRestModel                      # noqa: E402
from svb_models_asl_b200.svbcompat import nifti              # noqa: E402

TIS = [2.05, 2.3, 2.55, 2.8, 3.05, 3.3]
OPTIONS = {"tau": 1.8, "t1b": 1.6, "casl": True, "repeats": 1, "t1": 1.3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=10)
    ap.add_argument("--noise", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--outdir", default=".")
    a = ap.parse_args()
    n = a.side ** 3
    rng = np.random.default_rng(a.seed)
    model = AslRestModel(DataModel(np.zeros((1, len(TIS)), dtype=np.float32)), tis=TIS, **OPTIONS)
    ftiss = rng.uniform(1.0, 20.0, size=n)
    delttiss = rng.uniform(0.6, 2.5, size=n)
    params = np.zeros((2, n, 1), dtype=np.float32)
    params[0, :, 0], params[1, :, 0] = ftiss, delttiss
    tpts = np.tile(np.asarray(TIS, dtype=np.float32), (n, 1))
    sig = model.ievaluate(params, tpts)[:, 0, :]
    sig = rng.normal(sig, a.noise) if a.noise > 0 else sig
    shape = (a.side,) * 3
    os.makedirs(a.outdir, exist_ok=True)
    nifti.save(ftiss.reshape(shape).astype(np.float32), os.path.join(a.outdir, "ftiss.nii.gz"))
    nifti.save(delttiss.reshape(shape).astype(np.float32), os.path.join(a.outdir, "delttiss.nii.gz"))
    nifti.save(sig.reshape(shape + (-1,)).astype(np.float32), os.path.join(a.outdir, "sig.nii.gz"))
    print("Generated %i instances of test data in %s" % (n, a.outdir))


if __name__ == "__main__":
    main()
