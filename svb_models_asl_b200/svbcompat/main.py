"""
``run()``: mirror of ``svb.main.run`` as the reference scripts call it
(/root/reference/scripts/asl_example.py:16,45):

    runtime, svb, training_history = run(data, model_name, outdir, mask=..., **options)

Engine options understood (asl_example.py:29-41): learning_rate, batch_size, sample_size, epochs, log_stream,
save_mean, save_var, save_std, save_param_history, save_cost, save_cost_history, save_model_fit, save_log,
save_noise, save_runtime, force_num_latent_loss; everything else goes to the model (unknown keys are ignored,
as in svb).  Outputs: <outdir>/mean_<param>.nii.gz etc. shaped like the input volume (asl_example.py:47-54).
"""
import logging
import os
import sys

import numpy as np

from . import nifti
from .data import DataModel
from .fit import SvbFit


def run(data, model_name, output, mask=None, **kwargs):
    from ..plugin import get_model_class
    log = logging.getLogger("svb")
    stream = kwargs.get("log_stream", None)
    data_model = DataModel(data, mask, **kwargs)
    fwd_model = get_model_class(model_name)(data_model, **kwargs)
    tpts = fwd_model.tpts()
    svb = SvbFit(data_model, fwd_model, **kwargs)
    if stream is not None and svb.rank == 0:
        stream.write("Model: %s\n" % str(fwd_model))
        for p in svb.params:
            stream.write(" - %s: prior %s (%s), posterior %s\n" % (p.name, p.prior_dist, p.prior_type, p.post_dist))
    train_keys = ("batch_size", "epochs", "learning_rate", "sample_size", "display_step", "iters_per_launch")
    train_args = {k: kwargs[k] for k in train_keys if k in kwargs}
    rest = {k: v for k, v in kwargs.items() if k not in train_keys}
    history = svb.train(tpts, data_model.data_flattened, **train_args, **rest)
    runtime = svb.runtime
    log.info("DONE: %.3fs", runtime)

    means, variances = svb.model_moments()              # [P', n_local]
    means, variances = svb.gather(means.T), svb.gather(variances.T)       # [W, P']
    fit = svb.gather(svb.model_fit()) if kwargs.get("save_model_fit", False) else None
    final_cost = None
    if kwargs.get("save_cost", False) and "voxel_cost" not in history:
        # save_cost without save_cost_history: one evaluation of the per-voxel cost at the final posterior
        final_cost = svb.gather(svb.full_cost().cpu().numpy())
    if svb.world > 1:
        for key in ("voxel_cost", "params"):
            if key in history:
                history[key] = svb.gather(history[key])
    svb.close()
    if svb.rank != 0:
        return runtime, svb, history
    os.makedirs(output, exist_ok=True)

    def save(values, name):
        nifti.save(data_model.nifti_image(values), os.path.join(output, name + ".nii.gz"))

    for idx, param in enumerate(svb.params):
        is_noise = idx == len(svb.params) - 1
        if is_noise and not kwargs.get("save_noise", False):
            continue
        if kwargs.get("save_mean", False):
            save(means[:, idx], "mean_%s" % param.name)
        if kwargs.get("save_var", False):
            save(variances[:, idx], "var_%s" % param.name)
        if kwargs.get("save_std", False):
            save(np.sqrt(variances[:, idx]), "std_%s" % param.name)
        if kwargs.get("save_param_history", False) and "params" in history:
            save(history["params"][:, :, idx], "mean_%s_history" % param.name)
    if kwargs.get("save_cost", False) and "voxel_cost" in history:
        save(history["voxel_cost"][:, -1], "cost")
    elif final_cost is not None:
        save(final_cost, "cost")
    if kwargs.get("save_cost_history", False) and "voxel_cost" in history:
        save(history["voxel_cost"], "cost_history")
    if fit is not None:
        save(fit, "modelfit")
    if kwargs.get("save_runtime", False):
        with open(os.path.join(output, "runtime"), "w") as f:
            f.write("%f\n" % runtime)
    if kwargs.get("save_log", False):
        with open(os.path.join(output, "logfile"), "w") as f:
            f.write("model: %s\nruntime: %f s\nfinal mean cost: %f\nskipped non-finite updates: %i\n"
                    % (fwd_model, runtime, history["mean_cost"][-1], history.get("nan_skips", 0)))
    return runtime, svb, history


def main():
    """Minimal command line: svb --data <nii> --mask <nii> --model aslrest --output <dir> [--key value ...]"""
    args = sys.argv[1:]
    opts = {}
    key = None
    for a in args:
        if a.startswith("--"):
            key = a[2:].replace("-", "_")
            opts[key] = True
        elif key:
            try:
                opts[key] = float(a) if "." in a or "e" in a.lower() else int(a)
            except ValueError:
                opts[key] = [float(x) for x in a.split(",")] if "," in a else a
            key = None
    data, model, output = opts.pop("data"), opts.pop("model"), opts.pop("output", "svb_out")
    run(data, model, output, mask=opts.pop("mask", None), log_stream=sys.stdout, **opts)
