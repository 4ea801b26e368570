"""GPU timeline of one chunk-partitioned long-series step (torch.profiler, CUPTI): kernel count, busy time,
idle gaps.  usage: python tools/timeline_long.py [n] [ell] [sub]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cyclic-gps_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from cyclic_gps import _native, distributed as D
from cyclic_gps.synth import gaps_for_rows, leg_params, leg_precision_rows

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
ell = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sub = int(sys.argv[3]) if len(sys.argv) > 3 else None
dev = torch.device("cuda", 0)
_native.load()
plan = D.make_plan(n, 1, sub=sub)
G, Bm, LLT = leg_params(ell, seed=0, device=dev)
R, Oprev = leg_precision_rows(gaps_for_rows(0, n, n, seed=7, device=dev), G, Bm, LLT, torch.float32)
x = torch.randn((n, ell), dtype=torch.float32, device=dev)
R.requires_grad_(True); Oprev.requires_grad_(True); x.requires_grad_(True)

def step():
    R.grad = Oprev.grad = x.grad = None
    mh, ld = D.chunked_mahal_and_det(R, Oprev, x, plan, 0)
    (-0.5 * (mh + ld)).backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"kernels+memops {len(evs)}  span {(t1 - t0) / 1e3:.3f} ms  busy {busy / 1e3:.3f} ms")
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    k = e.name[:60]
    agg[k][0] += 1
    agg[k][1] += (e.time_range.end - e.time_range.start)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{c:5d} {t / 1e3:9.3f} ms  {k}")
# gaps
gaps = []
end = evs[0].time_range.end
for e in evs[1:]:
    if e.time_range.start > end:
        gaps.append((e.time_range.start - end, e.name[:40]))
    end = max(end, e.time_range.end)
gaps.sort(reverse=True)
print("idle total %.3f ms in %d gaps; largest:" % (sum(g for g, _ in gaps) / 1e3, len(gaps)))
for g, nm in gaps[:12]:
    print(f"   {g:8.1f} us before {nm}")
