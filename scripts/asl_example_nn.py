#!/usr/bin/env python
"""
Fit the aslnn surrogate model to multi-PLD pCASL difference data - the reference's scripts/asl_example_nn.py with its
option dict unchanged (:23-43: 6 PLDs x 8 repeats, slicedt 0.0452, batch_size 6, lr 0.01, train_load) and the same
read-back of mean_ftiss / mean_delttiss (:47-48); plotting needs matplotlib and is skipped when it is absent.

    python scripts/asl_example_nn.py /path/to/asldata_diff.nii.gz /path/to/asldata_mask.nii.gz
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svb.main import run  # noqa: E402

model = "aslnn"
outdir = "asl_example_out_nn"

# Inference options (reference: scripts/asl_example_nn.py:23-43)
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
    if not os.path.isdir(options["train_load"]):
        options["train_load"] = os.path.join(ROOT, "trained_data")
    runtime, svb, training_history = run(data, model, outdir, mask=mask, **options)
    from svb_models_asl_b200.svbcompat import nifti                    # nibabel stand-in (asl_example_nn.py:47-48)
    ftiss_img = nifti.load("%s/mean_ftiss.nii.gz" % outdir).data
    delttiss_img = nifti.load("%s/mean_delttiss.nii.gz" % outdir).data
    print("runtime %.2f s; mean ftiss %.3f, mean delttiss %.3f over the mask"
          % (runtime, ftiss_img[ftiss_img != 0].mean(), delttiss_img[delttiss_img != 0].mean()))
    try:
        import matplotlib.pyplot as plt                                 # asl_example_nn.py:50-55
        plt.figure("F")
        plt.imshow(ftiss_img[:, :, 10])
        plt.figure("delt")
        plt.imshow(delttiss_img[:, :, 10])
        plt.show()
    except ImportError:
        pass
