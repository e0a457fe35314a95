import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time
import manifold_constrained_gaussian_process_inference_b200 as pkg
from tests import helpers as H
for n, every, beta in [(161, 4, (1., 1., 1.)), (321, 8, (1., 1., 1.)), (161, 4, (1., 1., 4.))]:
    t = np.linspace(0.0, 20.0, n); truth = H.fn_truth(t); rng = np.random.default_rng(11)
    y = np.full_like(truth, np.nan); y[::every] = truth[::every] + 0.2 * rng.normal(size=truth[::every].shape)
    cfg = dict(niterHmc=1500, burninRatio=0.5, bandSize=20, stepSizeFactor=0.005, phi=np.array([[2.0, 1.0], [1.5, 2.0]]), priorTemperature=list(beta),
               sigmaInit=np.array([0.2, 0.2]), nChains=128, nLeapfrog=40, seed=2, thetaInit=np.array([0.5, 0.5, 2.0]))
    t0 = time.time(); res = pkg.solve_magi(y, t, pkg.fn_system(), cfg); dt = time.time() - t0
    st = res["stats"]
    print(n, every, beta, "theta", res["theta"].mean(axis=(0, 1)).round(3), "sigma", res["sigma"].mean(axis=(0, 1)).round(3), "acc", np.median(st["accept_rate"]).round(2),
          "eps", np.median(st["step_size"]).round(4), "xerr", np.abs(res["x_mean"].mean(axis=0) - truth).max().round(3), "time %.1fs" % dt, "evals/s %.2e" % (st["grad_evals"] / dt))
