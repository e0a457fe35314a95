"""Convergence experiment for the BASELINE config 5 problem (FitzHugh-Nagumo, n = 201, 41 noisy observations per component):
the chains start where the reference's solve_magi starts them (linear interpolation of the observations for X, src/MagiJl.jl:351-410)
instead of the throughput workload's truth + white noise, and split R-hat / bulk ESS of (theta, sigma, lp) are computed over the chains.

    python tools/cfg5_convergence.py --chains 1024 --iters 4000 --leapfrog 50 [--depth 7] [--fit-phi] [--beta 1,1,1]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--iters", type=int, default=4000)
ap.add_argument("--leapfrog", type=int, default=50)
ap.add_argument("--depth", type=int, default=0, help="> 0: batched NUTS with this maximum tree depth")
ap.add_argument("--fit-phi", action="store_true", help="estimate phi / sigma from the data (GP marginal likelihood) instead of the workload's phi")
ap.add_argument("--beta", default="1,1,1")
ap.add_argument("--step", type=float, default=0.005)
ap.add_argument("--n", type=int, default=201)
ap.add_argument("--rhat-chains", type=int, default=256)
ap.add_argument("--diag", action="store_true", help="print sampler health statistics (step sizes, metric range, runaway chains)")
ap.add_argument("--obs-every", type=int, default=0, help="observe every k-th grid point (default: about 41 observations)")
args = ap.parse_args()

if args.n != 201 or args.obs_every:
    synthetic.CONFIGS["fn201"] = dict(synthetic.CONFIGS["fn201"], n=args.n, obs_every=args.obs_every or max(1, (args.n - 1) // 40))
w = synthetic.make_workload("fn201", 8)
cfg = dict(niterHmc=args.iters, burninRatio=0.5, bandSize=20, stepSizeFactor=args.step, priorTemperature=[float(x) for x in args.beta.split(",")],
           nChains=args.chains, nLeapfrog=args.leapfrog, maxTreeDepth=args.depth, seed=5, setupMode="stable", xChains=1,
           thetaInit=np.array([0.5, 0.5, 2.0]))
if not args.fit_phi:
    cfg["phi"] = w["phi"]
    cfg["sigmaInit"] = w["sigma_init"]
t0 = time.time()
res = pkg.solve_magi(w["yobs"], w["tvec"], pkg.fn_system(), cfg)
dt = time.time() - t0
st = res["stats"]
k = 3
draws = np.concatenate([res["theta"], res["sigma"], res["lp"][..., None]], axis=2)        # (S, chains, 6)
sub = draws[:, :: max(1, draws.shape[1] // args.rhat_chains)][:, :args.rhat_chains]
thin = max(1, sub.shape[0] // 500)
summ = pkg.diagnostics.summarize(sub[::thin], names=["a", "b", "c", "sigma1", "sigma2", "lp"])
xm = res["x_mean"].mean(axis=0)
if args.diag:
    q = lambda v: np.percentile(np.asarray(v, dtype=np.float64), [0, 1, 50, 99, 100]).tolist()
    err = np.abs(res["x_mean"] - w["truth"][None]).max(axis=(1, 2))
    im = np.asarray(st.get("inv_metric", [np.nan]))
    print(json.dumps({"diag": True, "step_pct": q(st["step_size"]), "accept_pct": q(st["accept_rate"]), "divergences_pct": q(st.get("divergences", [0])),
                      "x_err_per_chain_pct": q(err), "chains_x_err_gt_5": int((err > 5).sum()), "inv_metric_pct": q(im[np.isfinite(im)]) if np.isfinite(im).any() else None,
                      "lp_last_pct": q(res["lp"][-1]), "stats_keys": sorted(st.keys())}))
print(json.dumps({"n": args.n, "n_obs": int(np.isfinite(w["yobs"][:, 0]).sum()), "chains": args.chains, "iters": args.iters, "leapfrog": args.leapfrog, "depth": args.depth, "fit_phi": args.fit_phi, "beta": args.beta,
                  "phi": np.asarray(res["phi"]).round(3).tolist(), "seconds": round(dt, 1), "grad_evals_per_s": float(st["grad_evals"] / dt),
                  "accept": float(np.median(st["accept_rate"])), "step": float(np.median(st["step_size"])),
                  "theta_mean": draws[..., :3].mean(axis=(0, 1)).round(4).tolist(), "theta_sd_between_chains": draws[..., :3].mean(axis=0).std(axis=0).round(4).tolist(),
                  "theta_sd_within": draws[..., :3].std(axis=0).mean(axis=0).round(4).tolist(),
                  "sigma_mean": draws[..., 3:5].mean(axis=(0, 1)).round(4).tolist(), "x_max_err": float(np.abs(xm - w["truth"]).max()),
                  "rhat": [round(r["rhat"], 3) for r in summ], "ess_bulk": [round(r["ess_bulk"], 1) for r in summ],
                  "rhat_chains": int(sub.shape[1]), "draws_per_chain_used": int(sub[::thin].shape[0])}))
