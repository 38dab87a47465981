"""cuBLAS DGEMM rate on this box (FP64 roofline denominator next to the DMMA microbenchmark)."""
import torch
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b.T
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); c = a @ b.T; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("cuBLAS DGEMM NT n=%d: %.2f ms  %.2f TFLOP/s" % (n, best, 2.0 * n ** 3 / best * 1e-9), flush=True)
