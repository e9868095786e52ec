"""cuBLAS DGEMM throughput on this GPU (library reference point for the FP64 roofline)."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 0.0
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
print(json.dumps({"cublas_dgemm_8192_tflops": round(best, 2)}))
