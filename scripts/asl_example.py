#!/usr/bin/env python
"""
Fit the aslrest model to multi-PLD pCASL difference data - the counterpart of the reference's
scripts/asl_example.py with its option dict unchanged (:24-42: 6 PLDs x 8 repeats, slicedt 0.0452, batch_size 6,
lr 0.01; plotting omitted).  Point it at the reference's own example data:

    python scripts/asl_example.py /path/to/asldata_diff.nii.gz /path/to/asldata_mask.nii.gz
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svb.main import run  # noqa: E402

model = "aslrest"
outdir = "asl_example_out"
# Inference options (reference: scripts/asl_example.py:24-42)
options = {
    "tau" : 1.8,
    "casl" : True,
    "plds" : [0.25, 0.5, 0.75, 1.0, 1.25, 1.5],
    "repeats" : [8],
    "slicedt" : 0.0452,
    "learning_rate" : 0.01,
    "batch_size" : 6,
    "sample_size" : 10,
    "epochs" : 500,
    "log_stream" : sys.stdout,
    "save_mean" : True,
    "save_var" : True,
    "save_param_history" : True,
    "save_cost" : True,
    "save_cost_history" : True,
    "save_model_fit" : True,
    "save_log" : True,
    "force_num_latent_loss" : True,
    "train_load" : "trained_data",
}

if __name__ == "__main__":
    data = sys.argv[1] if len(sys.argv) > 1 else "asldata_diff.nii.gz"
    mask = sys.argv[2] if len(sys.argv) > 2 else "asldata_mask.nii.gz"
    runtime, svb, training_history = run(data, model, outdir, mask=mask, **options)
    print("runtime %.2f s, final mean cost %.4f" % (runtime, training_history["mean_cost"][-1]))
