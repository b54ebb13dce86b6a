#!/usr/bin/env python
"""
Fit the aslrest model to simulated multi-PLD pCASL data - the counterpart of the reference's
scripts/asl_example_sim.py (same option dict, :23-40; plotting omitted) driving the B200 engine.

    python scripts/gen_test_data.py            # writes sig.nii.gz (+ ftiss/delttiss ground truth)
    python scripts/asl_example_sim.py [sig.nii.gz]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svb.main import run  # noqa: E402

model = "aslrest"
outdir = "asl_example_sim_out"
options = {
    "tau": 1.8, "casl": True, "plds": [0.25, 0.5, 0.75, 1.0, 1.25, 1.5], "repeats": [1],
    "learning_rate": 0.05, "sample_size": 10, "epochs": 5000, "log_stream": sys.stdout, "display_step": 500,
    "save_mean": True, "save_var": True, "save_param_history": False, "save_cost": True,
    "save_cost_history": False, "save_model_fit": True, "save_log": True, "force_num_latent_loss": True,
}

if __name__ == "__main__":
    data = sys.argv[1] if len(sys.argv) > 1 else "sig.nii.gz"
    runtime, svb, training_history = run(data, model, outdir, **options)
    print("runtime %.2f s, final mean cost %.4f" % (runtime, training_history["mean_cost"][-1]))
