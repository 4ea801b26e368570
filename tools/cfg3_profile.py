"""Where the CO2-shaped training step (configs[2]: n = 502, rank 16, fp64) goes: torch profiler table of one step.
usage (GPU box): python tools/cfg3_profile.py [rank] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    sys.path.insert(0, p)
from cyclic_gps.models import LEGFamily  # noqa: E402

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 502
torch.manual_seed(3)
gaps = torch.ones(n - 1, dtype=torch.float64)
gaps[n // 2] = 240.0
ts = (torch.cat([torch.zeros(1, dtype=torch.float64), torch.cumsum(gaps, 0)]) / 12.0).cuda()
xs = torch.randn(n, 1, dtype=torch.float64).cuda()
model = LEGFamily(rank=rank, obs_dim=1, train=True, data_type=torch.float64).cuda()


def step():
    model.zero_grad(set_to_none=True)
    ll = model.log_likelihood(ts, xs)
    (-ll / n).backward()
    return ll


for _ in range(5):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    step()
b.record()
torch.cuda.synchronize()
print("train step ms", a.elapsed_time(b) / 20)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=50))
ev = prof.key_averages()
nk = sum(e.count for e in ev if e.device_type == torch.autograd.DeviceType.CUDA)
print("device kernels per step", nk / 5)

import cProfile  # noqa: E402
import pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)

# the same step through cyclic_gps.graphs.GraphedLogLikelihood (one CUDA-graph replay for the device part)
from cyclic_gps.graphs import GraphedLogLikelihood  # noqa: E402
runner = GraphedLogLikelihood(model, ts, xs)


def gstep():
    model.zero_grad(set_to_none=True)
    ll = runner()
    (-ll / n).backward()
    with torch.no_grad():
        model.R_params.add_(1e-6)
    return ll


for _ in range(5):
    gstep()
torch.cuda.synchronize()
a.record()
for _ in range(20):
    gstep()
b.record()
torch.cuda.synchronize()
print("graphed train step ms (parameters moving)", a.elapsed_time(b) / 20)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    gstep()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
