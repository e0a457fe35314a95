"""BASELINE config 5 flow: FN n=201 chains sharded over the ranks of one box (torchrun), on-device HMC per shard with no
hot-path collective, then ONE NCCL all-gather of the retained (theta, sigma, lp) draws and R-hat / ESS on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_hmc.py --chains 65536 --iters 40 --warmup 20
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic, distributed as D

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=65536)
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--warmup", type=int, default=20)
ap.add_argument("--leapfrog", type=int, default=10)
ap.add_argument("--n", type=int, default=201, help="discretisation points (201 = BASELINE config 5; smaller grids converge within a short run)")
ap.add_argument("--step", type=float, default=0.002)
args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
first, n_local = D.shard_chains(args.chains, rank, world)
# every rank draws the SAME global chain population and keeps its slice: results do not depend on the world size
if args.n != 201:
    synthetic.CONFIGS["fn201"] = dict(synthetic.CONFIGS["fn201"], n=args.n, obs_every=max(1, (args.n - 1) // 40))
work = synthetic.make_workload("fn201", args.chains, rank=0)
params = work["params"][first:first + n_local]
tg = pkg.MagiTarget.from_config(work["yobs"], work["tvec"], work["phi"], pkg.fn_system(), work["sigma_init"], bandsize=20, jitter=1e-6,
                                setup_mode="stable", device=local, max_chains=n_local)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
chain, st = pkg.run_hmc_sampler(tg, params, n_samples=args.iters, n_adapts=args.warmup, initial_step_size=args.step, n_leapfrog=args.leapfrog,
                                seed=20251018 + 5, chain_id_offset=first, keep_on_device=True, n_chains_total=args.chains,
                                window_allreduce=D.make_window_allreduce(tg) if world > 1 else None)
torch.cuda.synchronize()
t_sample = time.perf_counter() - t0
draws = D.device_draws_as_tensor(tg)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
e0.record()
full = D.allgather_draws(draws)
e1.record(); torch.cuda.synchronize()
t_gather_ms = e0.elapsed_time(e1)
tt = torch.tensor([t_sample, st["grad_evals"]], dtype=torch.float64, device="cuda")
if world > 1:
    tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
else:
    tmax = tsum = tt
if rank == 0:
    d = full.cpu().numpy()
    sub = d[:, :: max(1, d.shape[1] // 256)][:, :256]
    summ = pkg.diagnostics.summarize(sub, names=["a", "b", "c", "sigma1", "sigma2", "lp"])
    print(json.dumps({"n_gpus": world, "chains_total": args.chains, "draw_tensor": list(d.shape), "allgather_ms": round(t_gather_ms, 3),
                      "allgather_GB": round(d.nbytes / 1e9, 4), "sample_seconds_max": float(tmax[0]), "grad_evals_total": int(tsum[1]),
                      "grad_evals_per_s": float(tsum[1] / tmax[0]), "checksum": float(np.nansum(d[..., :3])),
                      "theta_mean": np.nanmean(d[..., :3], axis=(0, 1)).round(4).tolist(), "rhat": [round(r["rhat"], 3) for r in summ],
                      "ess_bulk_256chains": [round(r["ess_bulk"], 1) for r in summ]}))
if world > 1:
    dist.destroy_process_group()
