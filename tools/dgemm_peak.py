# Measures cuBLAS DGEMM FP64 throughput (burst and sustained) as the FP64 roofline denominator.
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = 40
e0.record()
for _ in range(reps): c = a @ b
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"test": "cublas_dgemm_8192", "burst_tflops": round(burst, 2), "sustained_tflops": round(sus, 2), "sustained_seconds": e0.elapsed_time(e1) / 1e3}))
