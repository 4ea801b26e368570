"""Development aid: the warp-per-node DMMA kernels (variant 4) against the lane-per-row kernels (variant 1) and the CPU
oracle, output by output, without stopping at the first mismatch.   python tools/mma_debug.py [ell ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from cyclic_gps import _native, cyclic_reduction as c  # noqa: E402
from oracle import cr_oracle as orc  # noqa: E402
from test_cr_gpu import leg_inputs  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.shape != b.shape:
        return f"SHAPE {tuple(a.shape)} vs {tuple(b.shape)}"
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def run(l, n, dtype, variant):
    R, O, x = leg_inputs(l, n, dtype, seed=7 * l + n)
    _native.VARIANT = variant
    out = {}
    try:
        (mm_, K, F, G), (Rn, On) = c.decompose_step(R.cuda(), O.cuda())
        out.update(K=K, F=F, G=G, Rn=Rn, On=On)
        Rr, Or, xr = [t.cuda().requires_grad_(True) for t in (R, O, x)]
        mm, dd = c.mahal_and_det(Rr, Or, xr)
        (1.25 * mm + 0.75 * dd).backward()
        dec = c.decompose(R.cuda(), O.cuda())
        w = c.solve(dec, x.cuda())
        Sd, So = c.inverse_blocks(dec)
        out.update(mahal=mm, logdet=dd, gR=Rr.grad, gO=Or.grad, gx=xr.grad, solve=w, Sd=Sd, So=So)
        torch.cuda.synchronize()
    finally:
        _native.VARIANT = 0
    return out, (R, O, x)


def main():
    ells = [int(a) for a in sys.argv[1:]] or [8, 16, 12, 24, 32, 13]
    for l in ells:
        for dtype in (torch.float64, torch.float32):
            for n in (2, 3, 7, 64, 301):
                try:
                    got, (R, O, x) = run(l, n, dtype, 4)
                except Exception as ex:  # noqa: BLE001
                    print(f"l={l} {dtype} n={n}: EXCEPTION {type(ex).__name__}: {ex}")
                    continue
                Rd, Od, xd = R.double(), O.double(), x.double()
                (m2, K2, F2, G2), (Rn2, On2) = orc.level_step(Rd, Od)
                d_o = orc.factor(Rd, Od)
                gR, gO, gx = orc.loglik_grads(Rd, Od, xd, 1.25, 0.75)
                sd, so = orc.selected_inverse(d_o)
                want = dict(K=K2, F=F2, G=G2, Rn=Rn2, On=On2, mahal=orc.mahal(d_o, xd), logdet=orc.logdet(d_o), gR=gR, gO=gO, gx=gx,
                            solve=orc.solve(d_o, xd), Sd=sd, So=so)
                errs = {k: rel(got[k], want[k]) for k in want}
                tol = 1e-10 if dtype == torch.float64 else 1e-4
                bad = {k: v for k, v in errs.items() if isinstance(v, str) or not (v <= tol)}
                print(f"l={l} {str(dtype)[6:]} n={n}: " + ("OK  max %.2e" % max(errs.values()) if not bad else "BAD " + str(bad)), flush=True)


if __name__ == "__main__":
    main()
