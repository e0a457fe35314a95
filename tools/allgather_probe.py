"""Times consecutive in-library all-gathers of the draw store (torchrun, N ranks): first call against the following ones."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic, distributed as Dm
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
chains = int(os.environ.get("CHAINS", 65536)); iters = int(os.environ.get("ITERS", 120))
first, n_local = Dm.shard_chains(chains, rank, world)
work = synthetic.make_workload("fn201", chains, rank=0)
tg = pkg.MagiTarget.from_config(work["yobs"], work["tvec"], work["phi"], pkg.fn_system(), work["sigma_init"], bandsize=20, jitter=1e-6, setup_mode="stable", device=local, max_chains=n_local)
Dm.init_device_comm(tg)
st = torch.cuda.current_stream().cuda_stream
pkg.run_hmc_sampler(tg, work["params"][first:first + n_local], n_samples=iters, n_adapts=0, initial_step_size=0.002, n_leapfrog=1, seed=1, chain_id_offset=first,
                    keep_on_device=True, n_chains_total=chains, stream=st)
ts = []
for i in range(5):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); full = Dm.allgather_draws_device(tg, stream=st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = torch.tensor(ts, dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
# torch's own all_gather_into_tensor on the same data, for reference
d = Dm.device_draws_as_tensor(tg).contiguous(); out = torch.empty((world,) + tuple(d.shape), dtype=d.dtype, device=d.device)
tt = []
for i in range(3):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dist.all_gather_into_tensor(out, d); e1.record(); torch.cuda.synchronize(); tt.append(e0.elapsed_time(e1))
if rank == 0:
    print(json.dumps({"world": world, "bytes_total": int(full.numel() * 8), "library_allgather_ms": [round(float(x), 3) for x in t.cpu()], "torch_allgather_ms": [round(x, 3) for x in tt]}))
dist.destroy_process_group()
