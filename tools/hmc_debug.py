import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import manifold_constrained_gaussian_process_inference_b200 as pkg
from tests import helpers as H
t = np.arange(0.0, 5.0 + 1e-9, 0.5); truth = H.fn_truth(t); rng = np.random.default_rng(123); sig = np.array([0.25, 0.35])
y = truth + rng.normal(size=truth.shape) * sig
for niter, L in [(600, 25), (3000, 25)]:
    cfg = dict(niterHmc=niter, burninRatio=0.5, bandSize=20, stepSizeFactor=0.005, phi=np.array([[2.0, 1.0], [1.5, 2.0]]), sigmaInit=np.array([0.3, 0.3]), nChains=256, nLeapfrog=L, seed=1)
    res = pkg.solve_magi(y, t, pkg.fn_system(), cfg)
    st = res["stats"]
    print("niter", niter, "theta mean", res["theta"].mean(axis=(0, 1)).round(3), "sigma", res["sigma"].mean(axis=(0, 1)).round(3), "lp mean", res["lp"].mean().round(2),
          "acc", np.median(st["accept_rate"]).round(3), "eps", np.median(st["step_size"]).round(5), "div", st["n_divergent"].sum(), "minv range", st["inverse_metric"].min(), st["inverse_metric"].max())
    print("  lp first/last kept:", res["lp"][0].mean().round(2), res["lp"][-1].mean().round(2), " theta sd across chains of chain-means", res["theta"].mean(axis=0).std(axis=0).round(3))
    print("  rhat", [round(r["rhat"], 3) for r in pkg.diagnostics.summarize(res["theta"][:, :64])])
