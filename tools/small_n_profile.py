"""Host-side cost of small problems (configs[0]: n = 1000, l = 3, fp64): wall time per API call with and
without a trailing synchronize, and a cProfile of the hottest host functions.
usage: python tools/small_n_profile.py [n] [ell] [dtype]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cyclic-gps_b200"))
import torch
from cyclic_gps import cyclic_reduction as cr
from cyclic_gps.synth import leg_params, leg_precision_blocks

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000
ell = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dtype = getattr(torch, sys.argv[3]) if len(sys.argv) > 3 else torch.float64
dev = "cuda"
G, Bm, LLT = leg_params(ell, seed=1, device=dev)
gaps = -torch.log(torch.rand((1, n - 1), dtype=torch.float64, device=dev)) + 0.01
R, O = leg_precision_blocks(gaps, G, Bm, LLT, dtype)
R, O = R[0].contiguous(), O[0].contiguous()
x = torch.randn((n, ell), dtype=dtype, device=dev)
Rg, Og, xg = R.clone().requires_grad_(True), O.clone().requires_grad_(True), x.clone().requires_grad_(True)


def t_decompose():
    return cr.decompose(R, O)


dec = t_decompose()


def t_mahal_and_det():
    return cr.mahal_and_det(R, O, x)


def t_solve():
    return cr.solve(dec, x)


def t_inverse_blocks():
    return cr.inverse_blocks(dec)


def t_loglik_grad():
    Rg.grad = Og.grad = xg.grad = None
    m, d = cr.mahal_and_det(Rg, Og, xg)
    (m + d).backward()


def bench(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t_host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / reps
    return t_host * 1e6, t_all * 1e6


for f in (t_decompose, t_mahal_and_det, t_solve, t_inverse_blocks, t_loglik_grad):
    h, a = bench(f)
    print(f"{f.__name__[2:]:16s} host {h:8.1f} us/call   host+device {a:8.1f} us/call")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    t_loglik_grad()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
