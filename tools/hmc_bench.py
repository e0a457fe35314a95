"""Steady-state throughput of the on-device HMC sampler (gradient evaluations per second inside magi_hmc_run, warm-up and
allocation excluded): FN n=201, b=20, CHAINS chains, 10 leapfrog steps per transition."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic, _lib

L = _lib.lib()
for chains in [int(c) for c in os.environ.get("CHAINS", "4096,8192").split(",")]:
    w = synthetic.make_workload("fn201", chains)
    tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.fn_system(), w["sigma_init"], bandsize=20, jitter=1e-6, setup_mode="stable")
    p0 = np.ascontiguousarray(w["params"])
    _lib.check(L.magi_hmc_init(tg._h, chains, _lib.as_dp(p0), ctypes.c_ulonglong(7), 0.002, ctypes.c_longlong(0)))
    _lib.check(L.magi_hmc_run(tg._h, 60, 10, 1, 0.8, 0, None))          # warm-up (adapts step size and metric)
    ts = []
    for rep in range(5):
        t0 = time.perf_counter()
        _lib.check(L.magi_hmc_run(tg._h, 100, 10, 0, 0.8, 0, None))     # 100 transitions x 10 leapfrog steps, synchronises at the end
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print(json.dumps({"chains": chains, "fused": os.environ.get("MAGI_HMC_NO_FUSE") is None, "seconds_per_100_transitions": round(dt, 5),
                      "grad_evals_per_s": round(chains * 1000 / dt, 1), "us_per_leapfrog_step": round(dt / 1000 * 1e6, 2)}))
    tg.close()
