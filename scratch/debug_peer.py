import os, sys, numpy as np, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
td.init_process_group("nccl", device_id=torch.device("cuda", rank))
from svb import DataModel
from svb_models_asl import AslRestModel
from svb_models_asl_b200.svbcompat.fit import SvbFit
from svb_models_asl_b200.sharding import ShardPlan
def log(*a): print("[r%d]" % rank, *a, flush=True)
rng = np.random.default_rng(0)
vol = rng.normal(5, 1, (12, 10, 8, 6)).astype(np.float32)
dm = DataModel(vol)
model = AslRestModel(dm, tau=1.8, casl=True, plds=[0.25, 0.5, 0.75, 1.0, 1.25, 1.5], repeats=[1],
                     param_overrides={"ftiss": {"prior_type": "M"}})
fit = SvbFit(dm, model)
fit._setup(model.tpts(), dm.data_flattened, None, 10, 0.05, epochs=50, force_num_latent_loss=True,
           param_overrides={"ftiss": {"prior_type": "M"}}, halo_mode="none")
f = fit.fused; plan = fit.plan
log("plan", plan.lo, plan.hi, plan.halo_lo, plan.halo_hi, plan.prev_halo_hi, plan.next_halo_lo, "ld", f.ld)
f.share_state_with_neighbours(plan)
for side, p in f.peers.items():
    log("peer", side, "ptrs", [hex(x) for x in p["ptrs"]], "ld", p["ld"], "shift", p["shift"])
log("mirror", f._mirror, "ranges", f.ranges)
torch.cuda.synchronize(); td.barrier()
# 1. plain peer write test through fill_eps into the neighbour's state_alt buffer halo? use a scratch: skip
f.step_dev = torch.tensor([0], device=f.dev, dtype=torch.int64)
f._capturing = True
f._fork_stream = torch.cuda.Stream(device=f.dev)
for it in range(3):
    log("iter", it, "start")
    f._record_iteration()
    torch.cuda.synchronize()
    log("iter", it, "kernels ok")
    f.state, f.state_alt = f.state_alt, f.state
    td.barrier()
log("done", float(f.state[0, plan.halo_lo]), float(f.log_ak[0]))
td.destroy_process_group()
