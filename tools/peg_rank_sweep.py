import os, sys, json, torch
ROOT = "/root/repo"
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from cyclic_gps.peg import peg_precision
from test_peg_gpu import _model_G
def timeit(fn, k=5):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
for dtype in (torch.float64, torch.float32):
    for l in (4, 5, 6, 7, 8, 9, 10, 12):
        B, n = 64, 10000
        G, shift = _model_G(l, l)
        gaps = (torch.rand((B, n - 1), dtype=torch.float64, device="cuda") + 0.05).to(dtype)
        cR = torch.randn((B, n, l, l), dtype=dtype, device="cuda"); cO = torch.randn((B, n - 1, l, l), dtype=dtype, device="cuda")
        cl = torch.randn(B, dtype=torch.float64, device="cuda")
        Gd, sd = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
        def fwd():
            with torch.no_grad(): return peg_precision(gaps, Gd, sd, logdet=True)
        def both():
            R, O, ld = peg_precision(gaps, Gd, sd, logdet=True)
            torch.autograd.grad((R, O, ld), (Gd, sd), (cR, cO, cl))
        f, fb = timeit(fwd), timeit(both)
        print(json.dumps({"l": l, "dtype": str(dtype), "gaps": B * (n - 1), "fwd_ms": round(f, 3), "bwd_ms": round(fb - f, 3), "ns_per_gap_fwd": round(f * 1e6 / (B * n), 2), "ns_per_gap_bwd": round((fb - f) * 1e6 / (B * n), 2)}), flush=True)
