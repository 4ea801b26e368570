"""Experiment behind bench.py's end-to-end pipeline (configs[1] on one GPU): does the precision builder of chunk c + 1, issued on
its own stream, overlap the cyclic reduction of chunk c?  The builder is bound by instruction issue (FMA), the level kernels by
memory latency, so the two can share an SM's cycles when their CTAs are resident together.
usage: python tools/e2e_overlap.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    sys.path.insert(0, p)
import bench  # noqa: E402
from cyclic_gps import cyclic_reduction as cr  # noqa: E402

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
hout = torch.empty((2, B), dtype=dtype, pin_memory=True)
d_ts, d_xs = torch.empty_like(ts), torch.empty_like(xs)
del gaps64, ts, xs
model = bench.bench_model(ell, dtype, dev, train=False)
_, shift = model._obs_terms()
copy_s, build_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def plan(first_frac, chunks):
    if chunks <= 1:
        return [B]
    first = max(1, B // first_frac)
    rest = chunks - 1
    return [first] + [(B - first) // rest + (1 if i < (B - first) % rest else 0) for i in range(rest)]


def make_step(sizes, overlap):
    bounds = [0]
    for z in sizes:
        bounds.append(bounds[-1] + z)
    nchunk = len(sizes)

    def step():
        main_s = torch.cuda.current_stream()
        copy_s.wait_stream(main_s)
        build_s.wait_stream(main_s)
        ready = []
        with torch.cuda.stream(copy_s):
            for c in range(nchunk):
                sl = slice(bounds[c], bounds[c + 1])
                d_ts[sl].copy_(h_ts[sl], non_blocking=True)
                d_xs[sl].copy_(h_xs[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_s)
                ready.append(ev)
        tot = torch.zeros((), dtype=torch.float64, device=dev)

        def build(c):
            sl = slice(bounds[c], bounds[c + 1])
            s = build_s if overlap else main_s
            s.wait_event(ready[c])
            with torch.cuda.stream(s), torch.no_grad():
                Rs, Os = model._precision_blocks(d_ts[sl], shift)
                v = model.compute_v(d_xs[sl])
                ev = torch.cuda.Event()
                ev.record(s)
            for t in (Rs, Os, v):
                t.record_stream(main_s)
            return Rs, Os, v, ev

        nxt = build(0)
        for c in range(nchunk):
            sl = slice(bounds[c], bounds[c + 1])
            Rs, Os, v, ev = nxt
            if overlap and c + 1 < nchunk:
                nxt = build(c + 1)                  # issued BEFORE the reduction of chunk c so that both are in flight
            main_s.wait_event(ev)
            Rs.requires_grad_(True); Os.requires_grad_(True); v.requires_grad_(True)
            mm, dd = cr.mahal_and_det(Rs, Os, v)
            ll = -0.5 * (mm.double().sum() + dd.double().sum())
            ll.backward()
            tot += ll.detach()
            hout[0, sl].copy_(mm.detach(), non_blocking=True)
            hout[1, sl].copy_(dd.detach(), non_blocking=True)
            if not overlap and c + 1 < nchunk:
                nxt = build(c + 1)
        main_s.synchronize()
        return float(tot)

    return step


def timeit(fn, k=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        v = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k, v


out = []
for first_frac, chunks, overlap in ((8, 2, False), (8, 2, True), (8, 3, True), (8, 4, True), (16, 5, True), (16, 9, True), (8, 4, False)):
    sizes = plan(first_frac, chunks)
    ms, v = timeit(make_step(sizes, overlap))
    out.append({"sizes": sizes, "builder_on_own_stream": overlap, "ms_per_step": ms, "loglik": v})
    print(json.dumps(out[-1]), flush=True)
