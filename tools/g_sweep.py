"""Chain-groups-per-block sweep (MAGI_FORCE_G) for small batches: which block shape fills the machine best."""
import json, os, sys, subprocess
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import json, os, sys
sys.path.insert(0, os.path.dirname(%r))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic
name, chains = sys.argv[1], int(sys.argv[2])
w = synthetic.make_workload(name, chains)
sysm = pkg.lv_system() if name.startswith("lv") else pkg.fn_system()
tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], sysm, w["sigma_init"], bandsize=20, jitter=1e-6, setup_mode="stable")
dev = torch.device("cuda")
p = torch.from_numpy(w["params"]).to(dev); g = torch.empty_like(p); ll = torch.empty(chains, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3): tg.logdensity_and_gradient_batched_dev(chains, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tg.logdensity_and_gradient_batched_dev(chains, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps({"workload": name, "chains": chains, "G": os.environ.get("MAGI_FORCE_G", "auto"), "ms": round(float(np.median(ts)), 4)}))
''' % here
for name, chains in [("lv1281", 2048), ("lv1281", 4096), ("fn201", 1024), ("fn201", 2048), ("fn201", 3072), ("fn201", 4096), ("fn201", 8192)]:
    for G in ("4", "2", "1"):
        env = dict(os.environ, MAGI_FORCE_G=G)
        r = subprocess.run([sys.executable, "-c", code, name, str(chains)], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
