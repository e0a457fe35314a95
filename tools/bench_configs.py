"""Measurements for BASELINE.json configs 3 and 4 (development tool; bench.py is the contract benchmark for config 2):
  cfg3: Lotka-Volterra n=1281, 2048 chains, banded (b=20) vs dense (b=n-1, FP64 DMMA GEMM path)
  cfg4: Lorenz-96 D=64, n=2001: device GP setup time (K3-K6) and achieved FP64 TFLOP/s, plus a 64-chain evaluation
Prints one JSON line per measurement."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic

def timed_eval(tg, params, reps=10):
    nch, P = params.shape
    dev = torch.device("cuda")
    p = torch.from_numpy(params).to(dev); g = torch.empty_like(p); ll = torch.empty(nch, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st)
    torch.cuda.synchronize()
    flush = torch.empty(200 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), ll.cpu().numpy()

which = sys.argv[1:] or ["cfg3", "cfg4"]
if "cfg3" in which:
    w = synthetic.make_workload("lv1281", 2048)
    n, D = w["n"], w["D"]
    for b, label in [(20, "banded b=20"), (n - 1, "dense b=n-1")]:
        t0 = time.perf_counter()
        tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.lv_system(), w["sigma_init"], bandsize=b, jitter=1e-6, setup_mode="stable")
        setup_s = time.perf_counter() - t0
        ms, ll = timed_eval(tg, w["params"])
        flops = synthetic.algorithmic_flops_per_eval(n, D, b)
        print(json.dumps({"config": "cfg3 LV n=1281 D=2 2048 chains, " + label, "ms": round(ms, 4), "evals_per_s": round(2048 / ms * 1e3, 1),
                          "algorithmic_TFLOPs": round(2048 * flops / ms * 1e-9, 2), "frac_of_fp64_dmma_peak_37.1": round(2048 * flops / ms * 1e-9 / 37.1, 3),
                          "setup_seconds_incl_alloc": round(setup_s, 3), "ll_finite": bool(np.all(np.isfinite(ll)))}))
        tg.close()
if "cfg4" in which:
    rng = np.random.default_rng(20251018 + 3)
    n, D = 2001, 64
    tvec = np.linspace(0.0, 20.0, n)
    phi = np.stack([rng.uniform(10, 20, D), rng.uniform(0.2, 0.4, D)])
    Y = np.full((n, D), np.nan); Y[::10] = 8.0 + rng.normal(size=(len(tvec[::10]), D))
    for mode in ("stable", "reference_order"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tg = pkg.MagiTarget.from_config(Y, tvec, phi, pkg.get_ode_system("lorenz96", D), np.full(D, 0.5), bandsize=20, jitter=1e-6, setup_mode=mode)
        dt = time.perf_counter() - t0
        flop = (5 if mode == "stable" else 6) * n ** 3 * D
        rep = [tg.setup_status(d) for d in range(D)]
        print(json.dumps({"config": "cfg4 Lorenz-96 D=64 n=2001 device GP setup (K3-K6), mode " + mode, "seconds_incl_alloc_and_14GB_memset": round(dt, 3),
                          "nominal_TFLOP": round(flop * 1e-12, 2), "TFLOPs": round(flop / dt * 1e-12, 2), "repaired_pivots_total": [int(sum(r[0] for r in rep)), int(sum(r[1] for r in rep))]}))
        if mode == "stable":
            P = n * D + 1 + D
            params = np.concatenate([8.0 + rng.normal(size=(64, n * D)), 8.0 + 0.1 * rng.normal(size=(64, 1)), np.log(0.5) + 0.1 * rng.normal(size=(64, D))], axis=1)
            ms, ll = timed_eval(tg, params, reps=3)
            print(json.dumps({"config": "cfg4 Lorenz-96 64 chains evaluation (GEMM path, band-truncated operators)", "ms": round(ms, 3), "ll_finite": bool(np.all(np.isfinite(ll)))}))
        tg.close()
