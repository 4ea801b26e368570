"""Warp-per-gap precision builder (ranks 9..32) against the same formulas in torch ops (what ranks > 8 used before):
forward + backward over one batch of gaps.  usage (GPU box): python tools/pegw_timing.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from cyclic_gps.peg import peg_precision, peg_precision_torch  # noqa: E402
from test_peg_gpu import _model_G  # noqa: E402


def timeit(fn, k=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


out = []
for (l, dtype, B, n) in ((16, torch.float64, 1, 502), (16, torch.float64, 1, 100000), (16, torch.float64, 8, 100000), (16, torch.float32, 8, 100000), (32, torch.float64, 4, 50000),
                         (12, torch.float64, 8, 100000), (24, torch.float32, 4, 100000)):
    G, shift = _model_G(l, l)
    gaps = (torch.rand((B, n - 1), dtype=torch.float64, device="cuda") + 0.05).to(dtype)
    cR = torch.randn((B, n, l, l), dtype=dtype, device="cuda")
    cO = torch.randn((B, n - 1, l, l), dtype=dtype, device="cuda")
    cl = torch.randn(B, dtype=torch.float64, device="cuda")
    Gd, sd = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
    Gt, st = G.to("cuda", dtype).requires_grad_(True), shift.to("cuda", dtype).requires_grad_(True)

    def dev_fwd():
        with torch.no_grad():
            return peg_precision(gaps, Gd, sd, logdet=True)

    def dev_both():
        R, O, ld = peg_precision(gaps, Gd, sd, logdet=True)
        torch.autograd.grad((R, O, ld), (Gd, sd), (cR, cO, cl))

    def torch_fwd():
        with torch.no_grad():
            return peg_precision_torch(gaps, Gt, st, logdet=True)

    def torch_both():
        R, O, ld = peg_precision_torch(gaps, Gt, st, logdet=True)
        torch.autograd.grad((R, O, ld), (Gt, st), (cR, cO, cl.to(dtype)))

    row = {"l": l, "dtype": str(dtype), "batch": B, "n": n, "kernel_fwd_ms": timeit(dev_fwd), "kernel_fwd_bwd_ms": timeit(dev_both)}
    row["torch_ops_fwd_ms"] = timeit(torch_fwd, 2)
    if B * n <= 100000:      # torch autograd through matrix_exp / linalg.solve faults (illegal memory access) at 8 x 1e5 gaps of 16 x 16
        row["torch_ops_fwd_bwd_ms"] = timeit(torch_both, 2)
    rows = B * n
    bytes_fwd = rows * (2 * l * l) * torch.empty((), dtype=dtype).element_size()
    row["fwd_GBs_written"] = bytes_fwd / (row["kernel_fwd_ms"] * 1e-3) / 1e9
    out.append(row)
    print(json.dumps(row), flush=True)
with open(os.path.join(ROOT, "gpurun_out", "r2_pegw_timing.json"), "w") as fh:
    json.dump(out, fh, indent=1)
