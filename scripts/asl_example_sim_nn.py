#!/usr/bin/env python
"""
Fit the aslnn surrogate model to simulated multi-PLD pCASL data - the reference's scripts/asl_example_sim_nn.py with its
option dict unchanged (:23-41); `train_load` points at this repository's trained_data/ (the reference does not ship
its weights; scripts/retrain_model.py regenerates them).

    python scripts/gen_test_data.py
    python scripts/asl_example_sim_nn.py [sig.nii.gz]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svb.main import run  # noqa: E402

model = "aslnn"
outdir = "asl_example_sim_nn_out"

# Inference options (reference: scripts/asl_example_sim_nn.py:23-41)
options = {
    "tau" : 1.8,
    "casl" : True,
    "plds" : [0.25, 0.5, 0.75, 1.0, 1.25, 1.5],
    "repeats" : [1],
    "learning_rate" : 0.05,
    "sample_size" : 10,
    "epochs" : 5000,
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
    data = sys.argv[1] if len(sys.argv) > 1 else "sig.nii.gz"
    if not os.path.isdir(options["train_load"]):
        options["train_load"] = os.path.join(ROOT, "trained_data")     # run from another directory
    runtime, svb, training_history = run(data, model, outdir, **options)
    print("runtime %.2f s, final mean cost %.4f" % (runtime, training_history["mean_cost"][-1]))
