"""BASELINE config 4 evaluation (Lorenz-96 D=64, n=2001, band 20, 64 chains) a few times: for launch lists / ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg
rng = np.random.default_rng(20251018 + 3)
n, D, nch = 2001, 64, int(os.environ.get("CHAINS", 64))
tvec = np.linspace(0.0, 20.0, n)
phi = np.stack([rng.uniform(10, 20, D), rng.uniform(0.2, 0.4, D)])
Y = np.full((n, D), np.nan); Y[::10] = 8.0 + rng.normal(size=(len(tvec[::10]), D))
tg = pkg.MagiTarget.from_config(Y, tvec, phi, pkg.get_ode_system("lorenz96", D), np.full(D, 0.5), bandsize=20, jitter=1e-6, setup_mode="stable")
prm = np.concatenate([8.0 + rng.normal(size=(nch, n * D)), 8.0 + 0.1 * rng.normal(size=(nch, 1)), np.log(0.5) + 0.1 * rng.normal(size=(nch, D))], axis=1)
p = torch.from_numpy(prm).cuda(); g = torch.empty_like(p); ll = torch.empty(nch, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st)
torch.cuda.synchronize()
ts = []
for _ in range(int(os.environ.get("REPS", 10))):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("ok", bool(torch.isfinite(ll).all()), "ms per evaluation: median %.4f min %.4f" % (float(np.median(ts)), min(ts)))
