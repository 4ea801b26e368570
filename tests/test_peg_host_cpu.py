"""Host side of the precision builder's backward pass (cyclic_gps/peg.py): the kernel returns the weighted sums
S[row] = sum_g E[row][g] gA_g^T (include/crb200.h); `_EigConsts.finish_expm_adjoint` turns them into the cotangent of G.
Checked here on the CPU against torch autograd through matrix_exp (what the reference differentiates, model_utils.py:12-29)."""
import pytest
import torch

from cyclic_gps.peg import _EigConsts


def _weighted_sums(consts, gaps, gA):
    """What crb200_peg_precision_bwd accumulates, restated with torch ops in fp64."""
    l = gA.shape[-1]
    c = -0.5 * gaps
    lam = torch.complex(consts.buf[:l], consts.buf[l:2 * l])
    S = torch.zeros((2 * l, l, l), dtype=torch.float64)
    gAT = gA.transpose(-1, -2)
    for m, (r0, r1, i0, i1) in enumerate(consts.rows):
        e = torch.exp(c.to(torch.complex128) * lam[m])
        S[r0] = (e.real[:, None, None] * gAT).sum(0)
        S[r1] = ((c * e.real)[:, None, None] * gAT).sum(0)
        if i0 >= 0:
            S[i0] = (e.imag[:, None, None] * gAT).sum(0)
            S[i1] = ((c * e.imag)[:, None, None] * gAT).sum(0)
    return S


@pytest.mark.parametrize("l,kind", [(1, "leg"), (2, "leg"), (3, "leg"), (5, "leg"), (8, "leg"), (4, "symmetric"), (6, "repeated")])
def test_finish_expm_adjoint_matches_autograd(l, kind):
    gen = torch.Generator().manual_seed(10 + l)
    N = torch.randn((l, l), generator=gen, dtype=torch.float64)
    R = torch.randn((l, l), generator=gen, dtype=torch.float64)
    if kind == "leg":                 # G = N N^T + R - R^T (models.py:127-135): complex pairs (+ one real eigenvalue for odd l)
        G = N @ N.T + R - R.T
    elif kind == "symmetric":         # all eigenvalues real
        G = N @ N.T + torch.eye(l, dtype=torch.float64)
    else:                             # repeated eigenvalues: the c-weighted sums carry the off-diagonal pairs too
        G = 0.7 * torch.eye(l, dtype=torch.float64)
    gaps = torch.rand(40, generator=gen, dtype=torch.float64) + 0.01
    gA = torch.randn((40, l, l), generator=gen, dtype=torch.float64)
    Gp = G.clone().requires_grad_(True)
    A = torch.matrix_exp(-0.5 * gaps[:, None, None] * Gp)
    (want,) = torch.autograd.grad((A * gA).sum(), Gp)
    consts = _EigConsts(G, torch.device("cpu"))
    assert consts.folded
    got = consts.finish_expm_adjoint(_weighted_sums(consts, gaps, gA))
    err = float((got - want).abs().max() / want.abs().max())
    assert err <= 1e-9, err
