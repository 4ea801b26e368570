"""CUDA-event timeline of one end-to-end step of bench.py (configs[1], one GPU): when each H2D copy lands, when each chunk's
builder / cyclic reduction starts and ends on the compute stream, and how long the host needs to queue everything.
usage: python tools/e2e_timeline.py [first_frac] [chunks]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    sys.path.insert(0, p)
import bench  # noqa: E402
from cyclic_gps import cyclic_reduction as cr  # noqa: E402

first_frac = int(sys.argv[1]) if len(sys.argv) > 1 else 8
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, n, ell, dtype, dev = 1024, 10000, 8, torch.float32, torch.device("cuda")
cr.EAGER_PD_CHECK = False
cr.RELEASE_FACTORS_AFTER_BACKWARD = True
gen = torch.Generator(device=dev).manual_seed(1)
gaps64 = -torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64, device=dev)) + 0.01
ts = torch.cat([torch.zeros((B, 1), dtype=torch.float64, device=dev), torch.cumsum(gaps64, 1)], 1)
xs = torch.randn((B, n, 1), generator=gen, dtype=dtype, device=dev)
h_ts = torch.empty(ts.shape, dtype=torch.float64, pin_memory=True).copy_(ts)
h_xs = torch.empty(xs.shape, dtype=dtype, pin_memory=True).copy_(xs)
hout = torch.empty((2, B), dtype=dtype, pin_memory=True)
d_ts, d_xs = torch.empty_like(ts), torch.empty_like(xs)
del gaps64, ts, xs
model = bench.bench_model(ell, dtype, dev, train=False)
_, shift = model._obs_terms()
copy_s = torch.cuda.Stream(device=dev)
first = max(1, B // first_frac)
rest = chunks - 1
sizes = [first] + [(B - first) // rest + (1 if i < (B - first) % rest else 0) for i in range(rest)]
bounds = [0]
for z in sizes:
    bounds.append(bounds[-1] + z)


def step(marks=None):
    main_s = torch.cuda.current_stream()
    ev = lambda name, s=None: marks is not None and marks.append((name, _rec(s or main_s), time.perf_counter()))
    copy_s.wait_stream(main_s)
    ev("start")
    ready = []
    with torch.cuda.stream(copy_s):
        for c in range(len(sizes)):
            sl = slice(bounds[c], bounds[c + 1])
            d_ts[sl].copy_(h_ts[sl], non_blocking=True)
            d_xs[sl].copy_(h_xs[sl], non_blocking=True)
            e = torch.cuda.Event()
            e.record(copy_s)
            ready.append(e)
            ev(f"h2d{c} landed", copy_s)
    tot = torch.zeros((), dtype=torch.float64, device=dev)
    for c in range(len(sizes)):
        sl = slice(bounds[c], bounds[c + 1])
        main_s.wait_event(ready[c])
        ev(f"chunk{c} may start")
        with torch.no_grad():
            Rs, Os = model._precision_blocks(d_ts[sl], shift)
            ev(f"chunk{c} builder done")
            v = model.compute_v(d_xs[sl])
        Rs.requires_grad_(True); Os.requires_grad_(True); v.requires_grad_(True)
        mm, dd = cr.mahal_and_det(Rs, Os, v)
        ev(f"chunk{c} cr fwd done")
        ll = -0.5 * (mm.double().sum() + dd.double().sum())
        ll.backward()
        ev(f"chunk{c} cr bwd done")
        tot += ll.detach()
        hout[0, sl].copy_(mm.detach(), non_blocking=True)
        hout[1, sl].copy_(dd.detach(), non_blocking=True)
    ev("queued")
    main_s.synchronize()
    return hout


def _rec(s):
    e = torch.cuda.Event(enable_timing=True)
    e.record(s)
    return e


for _ in range(3):
    step()
torch.cuda.synchronize()
marks = []
t0 = time.perf_counter()
step(marks)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
base = marks[0][1]
out = {"sizes": sizes, "wall_ms": wall * 1e3, "timeline": [{"what": nm, "gpu_ms": base.elapsed_time(e), "host_queued_ms": (t - marks[0][2]) * 1e3} for nm, e, t in marks]}
print(json.dumps(out, indent=1))
