"""GPU timeline of one configs[1]-shaped step (torch.profiler): kernel count, busy time, idle gaps.
usage: python tools/timeline_batch.py [batch] [n] [ell]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cyclic-gps_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from cyclic_gps import _native, cyclic_reduction as cr
from cyclic_gps.synth import leg_params, leg_precision_blocks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10000
ell = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
_native.load()
G, Bm, LLT = leg_params(ell, seed=0, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
gaps = -torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64, device=dev)) + 0.01
R, O = leg_precision_blocks(gaps, G, Bm, LLT, torch.float32)
x = torch.randn((B, n, ell), generator=gen, dtype=torch.float32, device=dev)
R.requires_grad_(True); O.requires_grad_(True); x.requires_grad_(True)

def step():
    R.grad = O.grad = x.grad = None
    mh, ld = cr.mahal_and_det(R, O, x)
    (-0.5 * (mh + ld)).sum().backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print("ms per step (events, 10 steps): %.3f" % (e0.elapsed_time(e1) / 10))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"kernels+memops {len(evs)}  span {(t1 - t0) / 1e3:.3f} ms  busy {busy / 1e3:.3f} ms")
gaps_ = []
end = evs[0].time_range.end
for e in evs[1:]:
    if e.time_range.start > end:
        gaps_.append((e.time_range.start - end, e.name[:50]))
    end = max(end, e.time_range.end)
gaps_.sort(reverse=True)
print("idle total %.3f ms in %d gaps; largest:" % (sum(g for g, _ in gaps_) / 1e3, len(gaps_)))
for g, nm in gaps_[:10]:
    print(f"   {g:8.1f} us before {nm}")
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    agg[e.name[:60]][0] += 1
    agg[e.name[:60]][1] += (e.time_range.end - e.time_range.start)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{c:5d} {t / 1e3:9.3f} ms  {k}")
