# Probe: vendor bar for FP64 GEMM on this box (cuBLAS via torch.matmul), same recipe the driver used for bf16.
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / best * 1e-9
    # sustained: back-to-back ~2s
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); cnt = 0
    e0.record()
    while time.time() - t0 < 2.0:
        torch.matmul(a, b); cnt += 1
        if cnt % 4 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out[f"dgemm_{n}_tflops_sustained"] = 2 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
n = 16384
a = torch.randn(n, n, dtype=torch.float64, device="cuda").tril_() + 4 * torch.eye(n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); x = torch.linalg.solve_triangular(a, b, upper=False); e1.record(); torch.cuda.synchronize()
    out["cublas_dtrsm_16384_tflops"] = n**3 / e0.elapsed_time(e1) * 1e-9
print(json.dumps(out))
