"""CPU: host logic of the chunk-partitioned long-series path (cyclic_gps.distributed) -- plan,
halo bookkeeping, boundary-system assembly, the one all-gather, the descent -- exercised with an
oracle-backed engine (tests/cpu_engine.py), in-process for world=1 and with gloo for world=2,
against the unchunked oracle."""
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

from helpers import assert_close  # noqa: E402
from oracle import cr_oracle as orc  # noqa: E402


def make_series(n, l, seed):
    g = torch.Generator().manual_seed(seed)
    G, B, LLT = orc.leg_params(l, seed=seed)
    gaps = -torch.log(torch.rand(n - 1, generator=g, dtype=torch.float64)) + 0.05
    if n > 1:
        R, O = orc.leg_posterior_precision(gaps, G, B, LLT)
    else:                                   # a single row has no gaps: any SPD block will do
        R, O = (2.0 * torch.eye(l, dtype=torch.float64) + B.T @ B).unsqueeze(0), torch.zeros((0, l, l), dtype=torch.float64)
    x = torch.randn((n, l), generator=g, dtype=torch.float64)
    Oprev = torch.cat([torch.full((1, l, l), 7.0, dtype=torch.float64), O], dim=0)   # entry 0 must be ignored
    return R, O, Oprev, x


def run_rank(rank, world, n, l, sub, seed, gm, gd, group=None):
    import cpu_engine
    from cyclic_gps import distributed as D
    R, O, Oprev, x = make_series(n, l, seed)
    plan = D.make_plan(n, world, sub=sub)
    lo, hi = plan.rows(rank)
    Rl = R[lo:hi].clone().requires_grad_(True)
    Ol = Oprev[lo:hi].clone().requires_grad_(True)
    xl = x[lo:hi].clone().requires_grad_(True)
    mh, ld = D.chunked_mahal_and_det(Rl, Ol, xl, plan, rank, group=group, engine=cpu_engine)
    (gm * mh + gd * ld).backward()
    dec = orc.factor(R, O)
    assert_close(mh, orc.mahal(dec, x), 1e-11, f"mahal rank{rank}")
    assert_close(ld, orc.logdet(dec), 1e-11, f"logdet rank{rank}")
    gR, gO, gx = orc.loglik_grads(R, O, x, gm, gd)
    gOprev = torch.cat([torch.zeros(1, l, l, dtype=torch.float64), gO], dim=0)
    if hi > lo:
        assert_close(Rl.grad, gR[lo:hi], 1e-9, f"gR rank{rank}")
        assert_close(Ol.grad, gOprev[lo:hi], 1e-9, f"gO rank{rank}")
        assert_close(xl.grad, gx[lo:hi], 1e-9, f"gx rank{rank}")


@pytest.mark.parametrize("n,l,sub", [(32, 2, 8), (37, 3, 8), (64, 1, 4), (9, 2, 8), (8, 2, 8), (50, 3, 16), (7, 2, 2),
                                       (5, 2, 8), (1, 2, 2), (3, 1, 4)])   # series shorter than one sub-chunk: tail only
def test_chunked_single_process(n, l, sub):
    run_rank(0, 1, n, l, sub, seed=n + l, gm=0.7, gd=-1.3)


def _worker(rank, world, port, cases):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for (n, l, sub) in cases:
            run_rank(rank, world, n, l, sub, seed=3 * n + l, gm=1.0, gd=0.5, group=None)
    finally:
        dist.destroy_process_group()


def test_chunked_two_ranks_gloo():
    cases = [(64, 2, 8), (77, 3, 8), (40, 2, 4), (19, 2, 8),
             (5, 2, 8), (1, 2, 2), (12, 2, 8)]      # more ranks than sub-chunks: a rank owns nothing / only the ragged tail
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, cases), nprocs=2, join=True)


def test_plan_properties():
    from cyclic_gps import distributed as D
    p = D.make_plan(10 ** 8, 8)
    assert p.sub & (p.sub - 1) == 0 and p.nsub == -(-10 ** 8 // p.sub) and p.nboundary == 10 ** 8 // p.sub
    counts = [b - a for a, b in p.bounds]
    assert sum(counts) == p.nsub and max(counts) - min(counts) <= 1 and min(counts) >= 12
    lo, hi = p.rows(7)
    assert hi == 10 ** 8 and p.tail_rows(7) == 10 ** 8 % p.sub
    q = D.make_plan(10 ** 8, 8, sub=1 << 20)
    assert q.nsub == 96 and q.nboundary == 95 and max(b - a for a, b in q.bounds) == 12
    # more ranks than sub-chunks with a ragged tail: exactly one rank owns the tail, the surplus ranks own nothing
    t = D.make_plan(20, 4, sub=8)
    assert [t.rows(r) for r in range(4)] == [(0, 8), (8, 16), (16, 20), (20, 20)]
    assert [t.full_subchunks(r) for r in range(4)] == [1, 1, 0, 0] and [t.tail_rows(r) for r in range(4)] == [0, 0, 4, 0]
    covered = 0
    for r in range(8):
        a, b = p.rows(r)
        assert a == covered
        covered = b
    assert covered == 10 ** 8
