"""Dense-mode (band = n-1) evaluation of BASELINE config 3 (LV n=1281, 2048 chains): timing helper for ncu launch lists."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic

w = synthetic.make_workload("lv1281", int(os.environ.get("CHAINS", 2048)))
n, D = w["n"], w["D"]
tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.lv_system(), w["sigma_init"], bandsize=n - 1, jitter=1e-6, setup_mode="stable")
params = w["params"]; nch = params.shape[0]
dev = torch.device("cuda")
p = torch.from_numpy(params).to(dev); g = torch.empty_like(p); ll = torch.empty(nch, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get("REPS", 10))
for _ in range(2): tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
flops = synthetic.algorithmic_flops_per_eval(n, D, n - 1)
print(json.dumps({"config": "LV n=1281 dense", "chains": nch, "ms": round(ms, 4), "TFLOPs": round(nch * flops / ms * 1e-9, 2), "frac": round(nch * flops / ms * 1e-9 / 37.1, 3)}))
