"""BASELINE config 1 shape (fn_example.jl: FN, n=397, band 20, beta=[1,1,5]): latency of the single-chain drop-in call
magi_logdensity_and_gradient (host buffers), the call the reference's own NUTS loop would make once per leapfrog step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import manifold_constrained_gaussian_process_inference_b200 as pkg
from tests import helpers as H

prob = H.make_problem(model="fn", n=397, b=20, n_chains=4, seed=1, obs_every=4, T=20.0, beta=(1.0, 1.0, 5.0))
tg = H.cuda_target(pkg, prob)
p = prob["params"][0].copy()
for _ in range(20): tg.logdensity_and_gradient(p)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); ll, g = tg.logdensity_and_gradient(p); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e6
print(json.dumps({"config": "cfg1 shape: FN n=397 b=20 beta=[1,1,5], single-chain magi_logdensity_and_gradient through the Python mirror (host buffers)",
                  "median_us": round(float(np.median(ts)), 1), "p10_us": round(float(np.percentile(ts, 10)), 1), "p90_us": round(float(np.percentile(ts, 90)), 1),
                  "P": int(tg.dimension())}))
