"""GPU: the chunk-partitioned long-series path on the real CUDA engine (left-halo kernels,
boundary system, descent) against the unchunked oracle; world=1 in process and world=2 as two
processes sharing cuda:0 over gloo (the all-gather payload is < 1 kB per rank, so the transport
does not matter for correctness; NCCL is used by bench.py on real multi-GPU runs)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import TOL, assert_close  # noqa: E402
from oracle import cr_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def _series(n, l, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    G, B, LLT = orc.leg_params(l, seed=seed)
    gaps = -torch.log(torch.rand(n - 1, generator=g, dtype=torch.float64)) + 0.05
    R, O = orc.leg_posterior_precision(gaps, G, B, LLT)
    x = torch.randn((n, l), generator=g, dtype=torch.float64)
    Oprev = torch.cat([torch.full((1, l, l), 3.0, dtype=torch.float64), O], dim=0)
    return R.to(dtype), O.to(dtype), Oprev.to(dtype), x.to(dtype)


def _run(rank, world, n, l, dtype, sub, group=None, variant=0):
    from cyclic_gps import _native, distributed as D
    _native.VARIANT = variant
    R, O, Oprev, x = _series(n, l, dtype, seed=n + l)
    plan = D.make_plan(n, world, sub=sub)
    lo, hi = plan.rows(rank)
    Rl = R[lo:hi].cuda().requires_grad_(True)
    Ol = Oprev[lo:hi].cuda().requires_grad_(True)
    xl = x[lo:hi].cuda().requires_grad_(True)
    mh, ld = D.chunked_mahal_and_det(Rl, Ol, xl, plan, rank, group=group)
    (0.8 * mh - 0.6 * ld).backward()
    tol = TOL[dtype]
    Rd, Od, xd = R.double(), O.double(), x.double()
    dec = orc.factor(Rd, Od)
    assert_close(mh, orc.mahal(dec, xd), tol, "mahal")
    assert_close(ld, orc.logdet(dec), tol, "logdet")
    gR, gO, gx = orc.loglik_grads(Rd, Od, xd, 0.8, -0.6)
    gOprev = torch.cat([torch.zeros(1, l, l, dtype=torch.float64), gO], dim=0)
    assert_close(Rl.grad, gR[lo:hi], tol, "gR")
    assert_close(Ol.grad, gOprev[lo:hi], tol, "gOprev")
    assert_close(xl.grad, gx[lo:hi], tol, "gx")
    _native.VARIANT = 0


@pytest.mark.parametrize("n,l,dtype,sub,variant", [
    (5000, 4, torch.float32, 256, 0), (4096, 8, torch.float32, 512, 0), (3001, 3, torch.float64, 128, 0),
    (2050, 2, torch.float64, 64, 0), (1000, 8, torch.float64, 64, 0), (777, 16, torch.float64, 32, 0),
    (3001, 3, torch.float64, 128, 1), (1500, 8, torch.float32, 128, 1), (1500, 8, torch.float32, 128, 3),
    (1000, 8, torch.float64, 64, 1), (2050, 4, torch.float64, 64, 3),
    (520, 24, torch.float32, 32, 0), (300, 13, torch.float64, 16, 0), (400, 32, torch.float64, 32, 4), (900, 8, torch.float32, 64, 4)])
def test_chunked_world1(n, l, dtype, sub, variant):
    _run(0, 1, n, l, dtype, sub, variant=variant)


def _worker(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        _run(rank, world, 6000, 4, torch.float32, 256)
        _run(rank, world, 2500, 3, torch.float64, 128)
    finally:
        dist.destroy_process_group()


def test_chunked_world2_shared_gpu():
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)


def test_long_series_full_size_properties():
    """n = 4e6, l = 4, fp32 (what fits a quick test; the bench runs n = 1e8): the chunked path must agree with the
    plain single-sweep path, the solve must satisfy J w = x, and gx = 2 w."""
    from cyclic_gps import cyclic_reduction as c, distributed as D
    from cyclic_gps.synth import gaps_for_rows, leg_params, leg_precision_rows
    from helpers import relerr
    n, l = 4_000_000, 4
    dev = torch.device("cuda")
    G, Bm, LLT = leg_params(l, seed=0, device=dev)
    R, Oprev = leg_precision_rows(gaps_for_rows(0, n, n, seed=5, device=dev), G, Bm, LLT, torch.float32)
    x = torch.randn((n, l), dtype=torch.float32, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    O = Oprev[1:].contiguous()
    mm, dd = c.mahal_and_det(R, O, x)
    plan = D.make_plan(n, 1, sub=1 << 16)
    Rl, Ol, xl = R.clone().requires_grad_(True), Oprev.clone().requires_grad_(True), x.clone().requires_grad_(True)
    mc, dc = D.chunked_mahal_and_det(Rl, Ol, xl, plan, 0)
    assert relerr(mc, mm) < 1e-5 and relerr(dc, dd) < 1e-5
    mc.backward()
    dec = c.decompose(R, O)
    w = c.solve(dec, x)
    assert relerr(xl.grad, 2 * w) < 1e-4
    Jw = torch.einsum("nij,nj->ni", R.double(), w.double())
    Jw[1:] += torch.einsum("nij,nj->ni", O.double(), w[:-1].double())
    Jw[:-1] += torch.einsum("nji,nj->ni", O.double(), w[1:].double())
    assert float((Jw - x.double()).abs().max() / x.abs().max()) < 1e-4
    assert relerr(mm, (x.double() * w.double()).sum()) < 1e-5


def test_chunked_not_positive_definite_is_reported():
    """Default: raised by the forward call, with or without autograd.  With distributed.DEFERRED_PD_CHECK the report of
    the forward sweeps is read asynchronously and raised by the backward pass (distributed._finish_check)."""
    from cyclic_gps import distributed as D
    from cyclic_gps._engine import NotPositiveDefiniteError
    n, l = 3000, 4
    R, O, Oprev, x = _series(n, l, torch.float32, seed=3)
    R[1234] = -R[1234]
    plan = D.make_plan(n, 1, sub=256)
    with pytest.raises(NotPositiveDefiniteError):
        D.chunked_mahal_and_det(R.cuda(), Oprev.cuda(), x.cuda(), plan, 0)
    with pytest.raises(NotPositiveDefiniteError):
        D.chunked_mahal_and_det(R.cuda().requires_grad_(True), Oprev.cuda(), x.cuda(), plan, 0)
    D.DEFERRED_PD_CHECK = True
    try:
        Rl = R.cuda().requires_grad_(True)
        mh, ld = D.chunked_mahal_and_det(Rl, Oprev.cuda(), x.cuda(), plan, 0)
        with pytest.raises(NotPositiveDefiniteError):
            (mh + ld).backward()
    finally:
        D.DEFERRED_PD_CHECK = False
