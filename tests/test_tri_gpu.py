"""Packed lower triangles (include/crb200.h `tri`): in the fused likelihood path the blocks that never leave the library
(D, the reduced diagonal blocks, Sigma_d of the inner levels) travel through HBM as packed lower triangles where the
kernels offer it (float32, ell = 8).  The results must not depend on the storage: packed vs full blocks, sweep entries
vs per-level entries, and both against the CPU oracle (fp32 tolerance of BASELINE.json: 1e-4)."""
import pytest
import torch

from helpers import assert_close
from oracle import cr_oracle as orc
from test_cr_gpu import cr, leg_inputs

pytestmark = pytest.mark.gpu


def _step(c, R, O, x, gm=1.25, gd=0.75):
    Rr, Or, xr = [t.clone().requires_grad_(True) for t in (R, O, x)]
    mm, dd = c.mahal_and_det(Rr, Or, xr)
    ((gm * mm).sum() + (gd * dd).sum()).backward()
    return mm.detach(), dd.detach(), Rr.grad, Or.grad, xr.grad


def test_tri_is_offered_for_the_headline_block_size():
    from cyclic_gps import _native
    assert _native.tri_stride(torch.float32, 8) == 36
    assert _native.tri_stride(torch.float64, 8) == 0 and _native.tri_stride(torch.float32, 4) == 0


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 61, 62, 63, 64, 65, 66, 127, 128, 129, 1000, 1985, 4097])
def test_packed_equals_full_and_oracle_single_series(n):
    from cyclic_gps import _native
    c = cr()
    R, O, x = leg_inputs(8, n, torch.float32, seed=n)
    Rc, Oc, xc = R.cuda(), O.cuda(), x.cuda()
    assert _native.TRI
    packed = _step(c, Rc, Oc, xc)
    _native.TRI = False
    try:
        full = _step(c, Rc, Oc, xc)
    finally:
        _native.TRI = True
    for a, b, name in zip(packed, full, ("mahal", "logdet", "gR", "gO", "gx")):
        assert_close(a, b, 2e-6, f"{name} packed vs full, n={n}")
    if n <= 1000:
        Rd, Od, xd = R.double(), O.double(), x.double()
        dec = orc.factor(Rd, Od)
        gR, gO, gx = orc.loglik_grads(Rd, Od, xd, 1.25, 0.75)
        for got, want, name in ((packed[0], orc.mahal(dec, xd), "mahal"), (packed[1], orc.logdet(dec), "logdet"),
                                (packed[2], gR, "gR"), (packed[3], gO, "gO"), (packed[4], gx, "gx")):
            assert_close(got, want, 1e-4, f"{name} vs oracle, n={n}")
        # the gradient of the diagonal blocks is symmetric (the packed path only ever reads and writes lower triangles inside)
        assert_close(packed[2], packed[2].transpose(-1, -2), 1e-6, "gR symmetric")


@pytest.mark.parametrize("B,n", [(3, 500), (40, 257), (301, 130), (700, 64), (9, 10000)])
def test_packed_equals_full_batched(B, n):
    """Small batches take the four-warp fused tail, large ones the single-warp fused tail: both carry the flag per level."""
    from cyclic_gps import _native
    c = cr()
    parts = [leg_inputs(8, n, torch.float32, seed=100 + s) for s in range(min(B, 4))]
    R = torch.stack([parts[s % len(parts)][0] for s in range(B)]).cuda()
    O = torch.stack([parts[s % len(parts)][1] for s in range(B)]).cuda()
    x = torch.randn((B, n, 8), generator=torch.Generator().manual_seed(B), dtype=torch.float32).cuda()
    packed = _step(c, R, O, x)
    _native.TRI = False
    try:
        full = _step(c, R, O, x)
    finally:
        _native.TRI = True
    for a, b, name in zip(packed, full, ("mahal", "logdet", "gR", "gO", "gx")):
        assert_close(a, b, 2e-6, f"{name} packed vs full, B={B} n={n}")


def test_packed_no_grad_and_per_level_entries():
    """Without autograd no factors are kept (the reduced blocks are still packed); with bench.py's launch tracer the
    level loop runs in Python over the per-level entries, which take the same flag."""
    from cyclic_gps import _native

    class Tracer:
        enabled = True
        launches = 0

        def begin(self, kind, dtype, ell, batch, m):
            Tracer.launches += 1
            return None

        def end(self, tok):
            return None

    c = cr()
    R, O, x = (t.cuda() for t in leg_inputs(8, 1500, torch.float32, seed=3))
    ref = _step(c, R, O, x)
    with torch.no_grad():
        mm, dd = c.mahal_and_det(R, O, x)
    assert_close(mm, ref[0], 1e-6, "mahal without factors")
    assert_close(dd, ref[1], 1e-6, "logdet without factors")
    _native.TRACE = Tracer()
    try:
        traced = _step(c, R, O, x)
    finally:
        _native.TRACE = None
    assert Tracer.launches > 0
    for a, b, name in zip(traced, ref, ("mahal", "logdet", "gR", "gO", "gx")):
        assert_close(a, b, 1e-6, f"{name} per-level vs sweep (packed)")


def test_exposed_factors_stay_full():
    """decompose() hands the factors to the caller: full lower-triangular blocks with exact zeros above the diagonal, as in the reference."""
    c = cr()
    R, O, x = (t.cuda() for t in leg_inputs(8, 300, torch.float32, seed=4))
    ms, Ds, Fs, Gs = c.decompose(R, O)
    dec_o = orc.factor(R.double().cpu(), O.double().cpu())
    for k, (a, b) in enumerate(zip(Ds, dec_o[1])):
        assert_close(a, b, 1e-4, f"D[{k}]")
        assert torch.equal(torch.triu(a, diagonal=1), torch.zeros_like(a))
