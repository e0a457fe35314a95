"""PCIe ceilings next to the host-buffer (e2e) call: pinned H2D / D2H bandwidth alone and full duplex, and the call itself."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg
from manifold_constrained_gaussian_process_inference_b200 import synthetic

dev = torch.device("cuda")
for mb in (3.3, 13.3, 100.0):
    nbytes = int(mb * 1e6) // 8 * 8
    h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory(); h2 = torch.empty_like(h).pin_memory()
    d = torch.empty(nbytes // 8, dtype=torch.float64, device=dev); d2 = torch.empty_like(d)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(fn, reps=20):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    t_h2d = run(lambda: d.copy_(h, non_blocking=True))
    t_d2h = run(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    t_both = run(both)
    print(json.dumps({"MB": mb, "h2d_GBs": round(nbytes / t_h2d * 1e-9, 1), "d2h_GBs": round(nbytes / t_d2h * 1e-9, 1), "duplex_each_GBs": round(nbytes / t_both * 1e-9, 1),
                      "h2d_us": round(t_h2d * 1e6, 1), "duplex_us": round(t_both * 1e6, 1)}))
from manifold_constrained_gaussian_process_inference_b200 import _lib
L = _lib.lib()
for chains in (4096, 16384):
    w = synthetic.make_workload("fn201", chains)
    tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.fn_system(), w["sigma_init"], bandsize=20, jitter=1e-6, setup_mode="stable")
    hp = torch.from_numpy(w["params"]).pin_memory(); hg = torch.empty_like(hp).pin_memory(); hl = torch.empty(chains, dtype=torch.float64).pin_memory()
    a, b, c = hp.numpy(), hl.numpy(), hg.numpy()
    def step(): _lib.check(L.magi_logdensity_and_gradient_batched(tg._h, chains, _lib.as_dp(a), _lib.as_dp(b), _lib.as_dp(c)))
    for _ in range(5): step()
    t0 = time.perf_counter()
    for _ in range(30): step()
    dt = (time.perf_counter() - t0) / 30
    print(json.dumps({"e2e_chains": chains, "ms_per_call": round(dt * 1e3, 4), "evals_per_s": round(chains / dt, 1), "MB_each_way": round(hp.numel() * 8e-6, 2),
                      "effective_GBs_each_way": round(hp.numel() * 8 / dt * 1e-9, 1), "env_chunks": os.environ.get("MAGI_E2E_CHUNKS", "")}))
    tg.close()
