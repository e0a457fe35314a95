"""Comparator timings for the device GP setup ("the comparator to beat"): the same Lorenz-96 problem
(D = 64 dimensions, n = 2001) through the vendor libraries -- cuSOLVER potrf / potri and cuBLAS DGEMM as torch.linalg exposes
them -- next to libmagi_b200's own hand-written path (blocked Cholesky, recursive-doubling triangular inverse, DMMA GEMMs).
A tool only: nothing under manifold_constrained_gaussian_process_inference_b200/ calls a vendor solver.
Prints one JSON line; `python tools/setup_comparators.py > profiles/setup_comparators_r02.json`."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import manifold_constrained_gaussian_process_inference_b200 as pkg

dev = torch.device("cuda")
rng = np.random.default_rng(20251018 + 3)
n, D, eps = 2001, 64, 1e-6
tvec = np.linspace(0.0, 20.0, n)
phi = np.stack([rng.uniform(10, 20, D), rng.uniform(0.2, 0.4, D)])
Y = np.full((n, D), np.nan); Y[::10] = 8.0 + rng.normal(size=(len(tvec[::10]), D))

def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), r

# the 64 covariance matrices and derivative matrices, built with torch (Matern-5/2, gaussian_process.jl:78-123)
t = torch.tensor(tvec, device=dev)
var = torch.tensor(phi[0], device=dev)[:, None, None]; ell = torch.tensor(phi[1], device=dev)[:, None, None]
dt = (t[:, None] - t[None, :])[None]
r = dt.abs() / ell
s5 = 5.0 ** 0.5
C = var * (1 + s5 * r + 5 * r * r / 3) * torch.exp(-s5 * r)
Cp = -torch.sign(dt) * var * torch.exp(-s5 * r) * (5 * dt.abs() / (3 * ell ** 2) + 5 * s5 * dt.abs() ** 2 / (3 * ell ** 3))
Cpp = (5 * var / (3 * ell ** 2)) * torch.exp(-s5 * r) * (1 + s5 * r - 5 * r * r)
I = torch.eye(n, device=dev, dtype=torch.float64)[None]
Cj = C + eps * I

out = {"problem": "Lorenz-96 GP setup, D=%d dimensions, n=%d, batched over the dimensions, FP64" % (D, n), "gpu": torch.cuda.get_device_name(0)}
t_potrf, L = timed(lambda: torch.linalg.cholesky(Cj))
t_potri, Cinv = timed(lambda: torch.cholesky_inverse(L))
t_gemm, m = timed(lambda: torch.bmm(Cp, Cinv))
t_gemm2, K = timed(lambda: Cpp - torch.bmm(m, Cp.transpose(1, 2)))
t_trsm, W = timed(lambda: torch.linalg.solve_triangular(L, Cp.transpose(1, 2), upper=False))
t_syrk, _ = timed(lambda: torch.bmm(W.transpose(1, 2), W))
fl = float(n) ** 3 * D
out["vendor"] = {
    "cusolver_potrf_batched_s": t_potrf, "potrf_TFLOPs": fl / 3 / t_potrf * 1e-12,
    "cusolver_potri_s": t_potri, "potri_TFLOPs": 2 * fl / 3 / t_potri * 1e-12,
    "cublas_dgemm_bmm_s": t_gemm, "dgemm_TFLOPs": 2 * fl / t_gemm * 1e-12,
    "cublas_dgemm_bmm_plus_sub_s": t_gemm2,
    "cublas_trsm_s": t_trsm, "trsm_TFLOPs": fl / t_trsm * 1e-12,
    "cublas_gram_bmm_s": t_syrk,
    # the reference-order recipe with vendor kernels: 2 potrf + 2 potri + 2 GEMM;  the stable recipe: 2 potrf + 2 potri + TRSM + Gram + TRSM-like GEMM
    "reference_order_total_s": 2 * t_potrf + 2 * t_potri + t_gemm + t_gemm2,
    "stable_total_s": 2 * t_potrf + 2 * t_potri + t_trsm + t_syrk + t_gemm,
}
del C, Cp, Cpp, Cj, L, Cinv, m, K, W
torch.cuda.empty_cache()
ours = {}
for mode in ("stable", "reference_order"):
    best = None
    for _ in range(2):
        tg = pkg.MagiTarget.from_config(Y, tvec, phi, pkg.get_ode_system("lorenz96", D), np.full(D, 0.5), bandsize=20, jitter=eps, setup_mode=mode)
        k_ms, a_ms = tg.setup_timing(); tg.close()
        best = k_ms if best is None else min(best, k_ms)
    ours[mode + "_kernel_s"] = best * 1e-3
out["libmagi_b200"] = ours
out["ratio_vendor_over_ours"] = {"stable": out["vendor"]["stable_total_s"] / ours["stable_kernel_s"],
                                 "reference_order": out["vendor"]["reference_order_total_s"] / ours["reference_order_kernel_s"]}
print(json.dumps(out))
