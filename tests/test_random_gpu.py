"""Randomised sweep over (n, ell, dtype, batch) against the CPU oracle: every public entry of the CR path, with sizes
chosen around the tile boundaries of the kernels (31 / 32 even nodes per tile), the level at which the deep levels of a
sweep are fused into one launch, and the batch size at which that fusion is switched off (two series per SM)."""
import random

import pytest
import torch

from helpers import TOL, assert_close
from oracle import cr_oracle as orc
from test_cr_gpu import cr, leg_inputs

pytestmark = pytest.mark.gpu


def _cases():
    rng = random.Random(20260101)
    edge_n = [1, 2, 3, 4, 5, 31, 32, 33, 61, 62, 63, 64, 65, 123, 124, 125, 127, 128, 129, 247, 248, 249, 495, 496, 497, 1000, 1985, 2047, 2049]
    cases = []
    for i in range(44):
        n = rng.choice(edge_n) if i % 2 == 0 else rng.randint(1, 2600)
        l = rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 4, 9, 10, 12, 16])
        dtype = rng.choice([torch.float32, torch.float64])
        batch = rng.choice([1, 1, 2, 5])
        cases.append((n, l, dtype, batch, i))
    cases.append((700, 8, torch.float32, 300, 100))      # more than two series per SM: no fused tail
    cases.append((700, 8, torch.float32, 290, 101))      # fused tail with an almost full wave of CTAs
    cases.append((333, 3, torch.float64, 301, 102))
    # large batches: the single-tile deep levels run as ONE launch of single-warp CTAs (more CTAs than one wave holds)
    cases.append((2049, 4, torch.float32, 310, 103))
    cases.append((129, 5, torch.float64, 400, 104))
    cases.append((1000, 2, torch.float32, 1300, 105))
    cases.append((63, 8, torch.float32, 1500, 106))
    return cases


@pytest.mark.parametrize("n,l,dtype,batch,seed", _cases())
def test_random_case_vs_oracle(n, l, dtype, batch, seed):
    c = cr()
    tol = TOL[dtype]
    check = sorted(set([0, batch - 1, batch // 2]))       # series compared with the oracle (all are computed)
    base = [leg_inputs(l, n, dtype, seed=1000 * seed + b) for b in range(min(batch, 6))]
    pick = lambda b: base[b % len(base)]
    R = torch.stack([pick(b)[0] for b in range(batch)]).cuda()
    O = torch.stack([pick(b)[1] for b in range(batch)]).cuda()
    x = torch.stack([pick(b)[2] for b in range(batch)]).cuda()
    Rr, Or, xr = R.clone().requires_grad_(True), O.clone().requires_grad_(True), x.clone().requires_grad_(True)
    mm, dd = c.mahal_and_det(Rr, Or, xr)
    (0.7 * mm.sum() - 1.3 * dd.sum()).backward()
    dec = c.decompose(R, O)
    w = c.solve(dec, x)
    Sd, So = c.inverse_blocks(dec)
    mh2 = c.mahal(dec, x)
    ld2 = c.det(dec)
    for b in check:
        Rb, Ob, xb = (t.double() for t in pick(b))
        d_o = orc.factor(Rb, Ob)
        assert_close(mm[b], orc.mahal(d_o, xb), tol, "mahal")
        assert_close(dd[b], orc.logdet(d_o), tol, "logdet")
        assert_close(mh2[b], orc.mahal(d_o, xb), tol, "mahal(decomp)")
        assert_close(ld2[b], orc.logdet(d_o), tol, "det(decomp)")
        assert_close(w[b], orc.solve(d_o, xb), tol, "solve")
        sd, so = orc.selected_inverse(d_o)
        assert_close(Sd[b], sd, tol, "Sig_diag")
        if n > 1:
            assert_close(So[b], so, tol, "Sig_off")
        gR, gO, gx = orc.loglik_grads(Rb, Ob, xb, 0.7, -1.3)
        assert_close(Rr.grad[b], gR, tol, "gR")
        if n > 1:
            assert_close(Or.grad[b], gO, tol, "gO")
        assert_close(xr.grad[b], gx, tol, "gx")
