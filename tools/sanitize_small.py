"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck): tiny sizes, forward + backward,
decompose / solve / inverse_blocks, the chunked path with halos, the precision builder.
usage: compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from cyclic_gps import _native, cyclic_reduction as c, distributed as D  # noqa: E402
from cyclic_gps.peg import peg_precision  # noqa: E402
from test_cr_gpu import leg_inputs  # noqa: E402

cases = [(3, 70, torch.float64, 0), (8, 70, torch.float32, 0), (8, 40, torch.float64, 0), (16, 23, torch.float64, 4), (13, 9, torch.float32, 4),
         (24, 11, torch.float64, 4), (32, 7, torch.float32, 4), (12, 20, torch.float32, 1)]
for (l, n, dtype, variant) in cases:
    R, O, x = leg_inputs(l, n, dtype, seed=l + n)
    _native.VARIANT = variant
    Rr, Or, xr = [t.cuda().requires_grad_(True) for t in (R, O, x)]
    mm, dd = c.mahal_and_det(Rr, Or, xr)
    (mm + dd).backward()
    dec = c.decompose(R.cuda(), O.cuda())
    w = c.solve(dec, x.cuda())
    Sd, So = c.inverse_blocks(dec)
    _native.VARIANT = 0
    torch.cuda.synchronize()
    print("ok", l, n, dtype, variant, float(mm), flush=True)
# halo kernels (chunked path, one rank)
for (l, n, dtype, sub) in ((4, 300, torch.float32, 32), (16, 70, torch.float64, 16)):
    R, O, x = leg_inputs(l, n, dtype, seed=5)
    Oprev = torch.cat([torch.zeros(1, l, l, dtype=dtype), O], 0)
    plan = D.make_plan(n, 1, sub=sub)
    Rl, Ol, xl = R.cuda().requires_grad_(True), Oprev.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    mh, ld = D.chunked_mahal_and_det(Rl, Ol, xl, plan, 0)
    (mh + ld).backward()
    torch.cuda.synchronize()
    print("ok chunked", l, n, flush=True)
# precision builder
for l, dtype in ((5, torch.float64), (8, torch.float32)):
    G = torch.eye(l, dtype=torch.float64) * 1.2 + torch.triu(torch.ones(l, l, dtype=torch.float64), 1) * 0.1 - torch.tril(torch.ones(l, l, dtype=torch.float64), -1) * 0.1
    G.requires_grad_(True)
    gaps = torch.rand((3, 70), dtype=torch.float64).add(0.05).to(dtype).cuda()
    Rb, Ob = peg_precision(gaps, G, torch.eye(l, dtype=torch.float64) * 0.3)
    (Rb.sum() + Ob.sum()).backward()
    torch.cuda.synchronize()
    print("ok peg", l, flush=True)
print("SANITIZE_RUN_DONE")
