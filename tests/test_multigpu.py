"""
Two ranks / two GPUs (NCCL over NVLink) against one GPU, through svb.main.run:
* voxel-wise priors: shards are independent, results equal the single-GPU fit voxel for voxel
  (the Philox stream is keyed on the GLOBAL voxel id);
* spatial prior: the halo holds the neighbour ranks' boundary SAMPLES of every iteration - stored by the drawing kernel
  straight into the neighbour's halo columns over NVLink peer memory ("peer"), or sent with ncclSend/Recv ("nccl") -
  and the log-ak gradient is all-reduced every iteration (peer-memory mailboxes, or NCCL: "peer+nccl" / "nccl").
Skipped unless the box has at least two GPUs (gpurun --gpus 2).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLDS = [0.25, 0.5, 0.75, 1.0, 1.25, 1.5]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _options(spatial, halo_mode="peer"):
    opts = {"halo_mode": halo_mode,"tau": 1.8, "casl": True, "plds": PLDS, "repeats": [1], "learning_rate": 0.05, "sample_size": 10,
            "epochs": 60, "save_mean": True, "force_num_latent_loss": True, "display_step": 0, "inferart": not spatial}
    if spatial:
        opts["param_overrides"] = {"ftiss": {"prior_type": "M"}}
    return opts


def _worker(rank, world, port, vol_path, out_dir, spatial, halo_mode):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as td
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from svb.main import run
    _rt, svb, hist = run(vol_path, "aslrest", out_dir, **_options(spatial, halo_mode))
    if rank == 0:
        np.save(os.path.join(out_dir, "mean_cost.npy"), hist["mean_cost"])
        if spatial:
            np.save(os.path.join(out_dir, "log_ak.npy"), svb.fused.log_ak.cpu().numpy())
    td.destroy_process_group()


@pytest.mark.parametrize("spatial,halo_mode", [(False, "peer"), (True, "peer"), (True, "peer+nccl"), (True, "nccl")])
def test_two_gpus_reproduce_one_gpu(tmp_path, spatial, halo_mode):
    """halo_mode "peer": boundary state stored straight into the neighbour's halo over NVLink peer memory by the step
    kernel and log-ak gradient all-reduced over peer-memory mailboxes (svbasl_hyper_step_peers), whole iteration
    replayed as a CUDA graph; "peer+nccl": the same with an NCCL all-reduce; "nccl": explicit send/recv exchange on a
    side stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from svb.main import run
    from svb_models_asl_b200.svbcompat import nifti
    from tests.test_fit_gpu import _sim_volume
    rng = np.random.default_rng(9)
    vol, _f, _d = _sim_volume((12, 10, 8), rng, noise=1.0, t1b=1.65)
    vol_path = str(tmp_path / "sig.nii.gz")
    nifti.save(vol, vol_path)
    one = str(tmp_path / "one")
    _rt, svb1, hist1 = run(vol_path, "aslrest", one, **_options(spatial))
    two = str(tmp_path / "two")
    os.makedirs(two, exist_ok=True)
    mp.spawn(_worker, args=(2, _free_port(), vol_path, two, spatial, halo_mode), nprocs=2, join=True)
    f1 = nifti.load(os.path.join(one, "mean_ftiss.nii.gz")).data
    f2 = nifti.load(os.path.join(two, "mean_ftiss.nii.gz")).data
    d1 = nifti.load(os.path.join(one, "mean_delttiss.nii.gz")).data
    d2 = nifti.load(os.path.join(two, "mean_delttiss.nii.gz")).data
    if spatial:
        # the only cross-rank arithmetic is the double-precision sum of the log-ak gradient
        np.testing.assert_allclose(np.load(os.path.join(two, "log_ak.npy")), svb1.fused.log_ak.cpu().numpy(), rtol=1e-5)
        np.testing.assert_allclose(f2, f1, rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(d2, d1, rtol=2e-4, atol=2e-4)
    else:
        np.testing.assert_array_equal(f2, f1)
        np.testing.assert_array_equal(d2, d1)
    np.testing.assert_allclose(np.load(os.path.join(two, "mean_cost.npy")), hist1["mean_cost"], rtol=1e-5)
