"""Library fp64 GEMM rate (cuBLAS via torch.matmul), same method as MEASURED_PEAKS.json's bf16 figure:
8192^3, best of 10 (burst) and back-to-back for ~3 s (sustained). Test/measurement tool only."""
import json, time, torch
n = 8192
a = torch.randn(n, n, device="cuda", dtype=torch.float64)
b = torch.randn(n, n, device="cuda", dtype=torch.float64)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 0; t0 = time.time(); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5):
        c = a @ b
    reps += 5
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * n**3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
# potrf via cuSOLVER for context
m = a @ a.T + n * torch.eye(n, device="cuda", dtype=torch.float64)
torch.linalg.cholesky(m); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); L = torch.linalg.cholesky(m); e1.record(); torch.cuda.synchronize()
potrf_ms = e0.elapsed_time(e1)
e0.record(); Minv = torch.cholesky_inverse(L); e1.record(); torch.cuda.synchronize()
potri_ms = e0.elapsed_time(e1)
print(json.dumps({"cublas_dgemm_tflops_burst": burst, "cublas_dgemm_tflops_sustained": sust,
                  "cusolver_potrf_8192_ms": potrf_ms, "cusolver_potri_8192_ms": potri_ms}))
