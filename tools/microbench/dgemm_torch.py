"""cuBLAS DGEMM throughput through torch (float64 matmul) - the measured FP64 roofline denominator."""
import torch, json, time
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n ** 3 / best * 1e-9
    # sustained: 3 seconds back to back
    t0 = time.time(); k = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        c = a @ b; k += 1
        if k % 8 == 0:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_tflops_sustained"] = 2 * n ** 3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))
