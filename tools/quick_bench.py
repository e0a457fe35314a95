"""Ad-hoc device-resident timing of the banded kernel with oracle-built band tables (development aid, not bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg
from tests import helpers as H

n, b = int(os.environ.get("N", 201)), int(os.environ.get("B", 20))
for nch in [int(x) for x in os.environ.get("CHAINS", "4096,16384,65536").split(",")]:
    prob = H.make_problem(n=n, T=20.0, b=b, n_chains=8, seed=1, obs_every=5)
    tg = H.cuda_target(pkg, prob)
    P = tg.dimension()
    base = torch.tensor(prob["params"], device="cuda")
    params = base.repeat((nch + 7) // 8, 1)[:nch].contiguous()
    params += 1e-3 * torch.randn_like(params)
    ll = torch.empty(nch, dtype=torch.float64, device="cuda"); grad = torch.empty_like(params)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): tg.logdensity_and_gradient_batched_dev(nch, params.data_ptr(), ll.data_ptr(), grad.data_ptr(), st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); tg.logdensity_and_gradient_batched_dev(nch, params.data_ptr(), ll.data_ptr(), grad.data_ptr(), st); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    flops = 16 * 2 * (n * (2 * b + 1) - b * (b + 1)) + n * (12 + 50)
    print(json.dumps({"n": n, "b": b, "chains": nch, "ms": round(ms, 4), "min_ms": round(min(ts), 4), "evals_per_s": round(nch / ms * 1e3, 1),
                      "GBs": round(nch * 8 * (2 * P + 1) / ms * 1e-6, 1), "useful_TF": round(nch * flops / ms * 1e-9, 2)}))
