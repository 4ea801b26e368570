"""Where the end-to-end step of bench.py goes (configs[1] on one GPU): CUDA-event timings of its pieces.
usage: python tools/e2e_breakdown.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    sys.path.insert(0, p)
import bench  # noqa: E402
from cyclic_gps import cyclic_reduction as cr  # noqa: E402
from cyclic_gps.peg import peg_precision  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, ell, dtype, dev = 10000, 8, torch.float32, torch.device("cuda")
cr.EAGER_PD_CHECK = False
cr.RELEASE_FACTORS_AFTER_BACKWARD = True
gen = torch.Generator(device=dev).manual_seed(1)
gaps64 = -torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64, device=dev)) + 0.01
ts = torch.cat([torch.zeros((B, 1), dtype=torch.float64, device=dev), torch.cumsum(gaps64, 1)], 1)
xs = torch.randn((B, n, 1), generator=gen, dtype=dtype, device=dev)
h_ts = torch.empty(ts.shape, dtype=torch.float64, pin_memory=True).copy_(ts)
h_xs = torch.empty(xs.shape, dtype=dtype, pin_memory=True).copy_(xs)
model = bench.bench_model(ell, dtype, dev, train=False)
tmodel = bench.bench_model(ell, dtype, dev, train=True)
_, shift = model._obs_terms()
gaps = gaps64.to(dtype)


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


out = {}
out["h2d_ts_xs_ms"] = timeit(lambda: (ts.copy_(h_ts, non_blocking=True), xs.copy_(h_xs, non_blocking=True)))
out["gaps_from_ts_ms"] = timeit(lambda: (ts[..., 1:] - ts[..., :-1]).to(dtype))
with torch.no_grad():
    out["builder_fwd_ms"] = timeit(lambda: peg_precision(gaps, model.G, shift))
    out["compute_v_ms"] = timeit(lambda: model.compute_v(xs))
    Rs, Os = peg_precision(gaps, model.G, shift)
    v = model.compute_v(xs)


def cr_step():
    R, O, x = Rs.detach().requires_grad_(True), Os.detach().requires_grad_(True), v.detach().requires_grad_(True)
    mm, dd = cr.mahal_and_det(R, O, x)
    (-0.5 * (mm.double().sum() + dd.double().sum())).backward()


out["cr_fwd_bwd_ms"] = timeit(cr_step)
Gp = model.G.clone().requires_grad_(True)
sp = shift.clone().requires_grad_(True)
R2, O2 = peg_precision(gaps, Gp, sp)
gR, gO = torch.randn_like(R2), torch.randn_like(O2)
out["builder_bwd_ms"] = timeit(lambda: torch.autograd.grad((R2, O2), (Gp, sp), (gR, gO), retain_graph=True))
del R2, O2, gR, gO


def train():
    for p in tmodel.parameters():
        p.grad = None
    tmodel.log_likelihood(ts, xs).sum().backward()


out["train_step_device_inputs_ms"] = timeit(train, 3)
print(json.dumps(out, indent=1))
