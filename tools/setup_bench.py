"""Device GP setup (K3-K6) of BASELINE config 4 (Lorenz-96, D=64, n=2001): wall time of repeated creates (the first one pays
the allocator), for an ncu launch list of where the time goes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg

n, D = int(os.environ.get("N", 2001)), int(os.environ.get("D", 64))
rng = np.random.default_rng(20251018 + 3)
tvec = np.linspace(0.0, 20.0, n)
phi = np.stack([rng.uniform(10, 20, D), rng.uniform(0.2, 0.4, D)])
Y = np.full((n, D), np.nan); Y[::10] = 8.0 + rng.normal(size=(len(tvec[::10]), D))
for mode in os.environ.get("MODES", "stable,reference_order").split(","):
    for rep in range(int(os.environ.get("REPS", 3))):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tg = pkg.MagiTarget.from_config(Y, tvec, phi, pkg.get_ode_system("lorenz96", D), np.full(D, 0.5), bandsize=20, jitter=1e-6, setup_mode=mode)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        flop = (5 if mode == "stable" else 6) * n ** 3 * D
        print(json.dumps({"mode": mode, "rep": rep, "seconds": round(dt, 4), "nominal_TFLOP": round(flop * 1e-12, 2), "TFLOPs": round(flop / dt * 1e-12, 2)}))
        tg.close()
