"""CPU oracle for the cyclic-reduction (CR) hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker / CPU
comparator.  The shipped path (``cyclic-gps_b200/``) never imports it and has no CPU
fallback.

What this is: an independent restatement, in plain torch-on-CPU tensor algebra, of the
algorithm in the reference's ``cyclic_gps/cyclic_reduction.py`` (block cyclic reduction =
block Cholesky of a symmetric positive-definite block-tridiagonal matrix in recursive
even/odd elimination order).  It uses the same ATen primitives the reference does
(batched Cholesky, triangular solve, matmul), so timing it on host cores is a fair
"port" of the reference CPU path, and autograd through it reproduces the reference's
backward.

Parity pinning: ``tests/golden/*.npz`` were produced by importing the UNMODIFIED
reference (``/root/reference``, through ``oracle/ref_shims``) in the build container with
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks every function
here against those vectors (which include the reference's own known-answer cases:
random block-bidiagonal LL^T, BAB and Schur-block matrices).  Unpinned: the
non-positive-definite jitter-retry path of gpytorch's ``psd_safe_cholesky``
(``cyclic_reduction.py:227,306,429``) -- this oracle raises instead.

Notation: level with m rows has e=ceil(m/2) even (eliminated) nodes,
o=floor(m/2) odd (surviving) nodes and g=floor((m-1)/2) "G" links.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

Tensor = torch.Tensor


class NotPositiveDefinite(RuntimeError):
    """A diagonal block met during elimination was not positive definite."""


def _chol(A: Tensor) -> Tensor:
    # reference: psd_safe_cholesky(...) at cyclic_reduction.py:227,306,429 -- on PD input
    # that is torch.linalg.cholesky_ex; the jitter ladder is an error path (unpinned).
    L, info = torch.linalg.cholesky_ex(A)
    if bool(torch.any(info)):
        raise NotPositiveDefinite(f"block {int(torch.nonzero(info)[0])} is not positive definite")
    return L


def _lsolve(L: Tensor, B: Tensor) -> Tensor:
    """L^{-1} B for lower-triangular L (batched)."""
    return torch.linalg.solve_triangular(L, B, upper=False)


def _ltsolve(L: Tensor, B: Tensor) -> Tensor:
    """L^{-T} B for lower-triangular L (batched)."""
    return torch.linalg.solve_triangular(L.mT, B, upper=True)


# --------------------------------------------------------------------------------------
# block-bidiagonal helper products (reference cyclic_reduction.py:15-200, Appendix A of
# SURVEY.md).  U has diagonal blocks F_j (odd node j x even node j) and super-diagonal
# blocks G_j (odd node j x even node j+1); "wide" = len(G)==len(F) (m odd).
# --------------------------------------------------------------------------------------
def bidiag_gram(F: Tensor, G: Tensor) -> Tuple[Tensor, Tensor]:
    """Tridiagonal blocks of U U^T: diag_j = F_j F_j^T + G_j G_j^T, low_j = F_{j+1} G_j^T.
    reference UU_T :15-37."""
    o, g = F.shape[0], G.shape[0]
    diag = F @ F.mT
    gg = G @ G.mT
    if g == o:
        diag = diag + gg
    else:
        diag = torch.cat([diag[:g] + gg, diag[g:]], dim=0)
    low = F[1:] @ G[: max(o - 1, 0)].mT
    return diag, low


def bidiag_mv(F: Tensor, G: Tensor, x: Tensor) -> Tensor:
    """U x : row j = F_j x_j + G_j x_{j+1}.  reference Ux :40-60."""
    o, g = F.shape[0], G.shape[0]
    out = torch.einsum("jab,jb->ja", F, x[:o])
    gx = torch.einsum("jab,jb->ja", G, x[1 : g + 1])
    if g == o:
        return out + gx
    return torch.cat([out[:g] + gx, out[g:]], dim=0)


def bidiag_tmv(F: Tensor, G: Tensor, x: Tensor) -> Tensor:
    """U^T x : row j = F_j^T x_j + G_{j-1}^T x_{j-1}; o+1 rows when wide, o rows otherwise.
    reference U_Tx :63-87."""
    o, g = F.shape[0], G.shape[0]
    ft = torch.einsum("jba,jb->ja", F, x[:o])
    gt = torch.einsum("jba,jb->ja", G, x[:g])
    rows = o + 1 if g == o else o
    out = x.new_zeros((rows, x.shape[1]))
    out = out + torch.cat([ft, x.new_zeros((rows - o, x.shape[1]))], dim=0)
    out = out + torch.cat([x.new_zeros((1, x.shape[1])), gt, x.new_zeros((rows - 1 - g, x.shape[1]))], dim=0)
    return out


def symtri_times_bidiag(Sd: Tensor, So: Tensor, F: Tensor, G: Tensor) -> Tuple[Tensor, Tensor]:
    """Diagonal and super-diagonal blocks of Sigma U for symmetric block-tridiagonal
    Sigma (Sd diagonal, So lower blocks).  reference SigU :90-136."""
    o, g = F.shape[0], G.shape[0]
    mid = Sd @ F
    if o > 1:
        mid = torch.cat([mid[:1], mid[1:] + So @ G[: o - 1]], dim=0)
    hi = Sd[:g] @ G
    k = min(g, o - 1)
    if k > 0:
        hi = torch.cat([hi[:k] + So[:k].mT @ F[1 : k + 1], hi[k:]], dim=0)
    return mid, hi


def bidiag_t_bidiag_diag(Fu: Tensor, Gu: Tensor, Fv: Tensor, Gv: Tensor) -> Tensor:
    """Diagonal blocks of U^T V for two bidiagonal matrices of the same shape.
    reference UtV_diags :139-178."""
    o, g = Fu.shape[0], Gu.shape[0]
    ff = Fu.mT @ Fv
    gg = Gu.mT @ Gv
    rows = o + 1 if g == o else o
    z = Fu.new_zeros((1,) + tuple(Fu.shape[1:]))
    a = torch.cat([ff] + ([z] if rows > o else []), dim=0)
    b = torch.cat([z, gg] + ([z] * (rows - 1 - g)), dim=0)
    return a + b


def interleave(a: Tensor, b: Tensor) -> Tensor:
    """out[0::2]=a, out[1::2]=b, leftover of the longer appended.  reference :181-200."""
    k = min(a.shape[0], b.shape[0])
    head = torch.stack([a[:k], b[:k]], dim=1).reshape((2 * k,) + tuple(a.shape[1:]))
    return torch.cat([head, a[k:], b[k:]], dim=0)


# --------------------------------------------------------------------------------------
# one level, the factorisation, the two half solves
# --------------------------------------------------------------------------------------
def level_step(R: Tensor, O: Tensor):
    """One CR level.  reference decompose_step :204-259.
    Returns (m, K, F, G), (R_next, O_next)."""
    m = R.shape[0]
    assert O.shape[0] == m - 1, "need one fewer off-diagonal block than diagonal blocks"  # :223
    o, g = m // 2, (m - 1) // 2
    K = _chol(R[0::2])                                   # :225-227
    F = _lsolve(K[:o], O[0::2].mT).mT                    # O_{2j} K_j^{-T}      :242-244
    G = _lsolve(K[1 : 1 + g], O[1::2]).mT                # O_{2j+1}^T K_{j+1}^{-T}  :246-248
    gram_d, gram_o = bidiag_gram(F, G)
    return (m, K, F, G), (R[1::2] - gram_d, -gram_o)     # :253-254


def factor(R: Tensor, O: Tensor):
    """Full CR factorisation.  reference decompose :288-309.
    Returns (ms int64 tensor, Ds, Fs, Gs)."""
    ms: List[int] = []
    Ds: List[Tensor] = []
    Fs: List[Tensor] = []
    Gs: List[Tensor] = []
    while R.shape[0] > 1:
        (m, K, F, G), (R, O) = level_step(R, O)
        ms.append(m)
        Ds.append(K)
        Fs.append(F)
        Gs.append(G)
    Ds.append(_chol(R))
    ms.append(1)
    return torch.tensor(ms, dtype=torch.int64), Ds, Fs, Gs


def forward_sub(decomp, y: Tensor) -> List[Tensor]:
    """L^{-1} T y in CR order, one tensor per level.  reference halfsolve :312-338."""
    ms, Ds, Fs, Gs = decomp
    out: List[Tensor] = []
    cur = y
    for k in range(len(Ds)):
        xk = _lsolve(Ds[k], cur[0::2].unsqueeze(-1)).squeeze(-1)
        out.append(xk)
        if cur.shape[0] > 1:
            cur = cur[1::2] - bidiag_mv(Fs[k], Gs[k], xk)
    return out


def backward_sub(decomp, ycrr: Sequence[Tensor]) -> Tensor:
    """T^T L^{-T} y, bottom-up with re-interleaving.  reference backhalfsolve :341-377."""
    ms, Ds, Fs, Gs = decomp
    w = _ltsolve(Ds[-1], ycrr[-1].unsqueeze(-1)).squeeze(-1)
    for k in range(len(Ds) - 2, -1, -1):
        rhs = ycrr[k] - bidiag_tmv(Fs[k], Gs[k], w)
        w_even = _ltsolve(Ds[k], rhs.unsqueeze(-1)).squeeze(-1)
        w = interleave(w_even, w)
    return w


def solve(decomp, y: Tensor) -> Tensor:
    """J^{-1} y.  reference solve :441-444."""
    return backward_sub(decomp, forward_sub(decomp, y))


def logdet(decomp) -> Tensor:
    """log|J| = 2 sum log diag(D).  reference det :447-458 (vectorised, same value)."""
    ms, Ds, Fs, Gs = decomp
    total = Ds[0].new_zeros(())
    for D in Ds:
        total = total + torch.log(torch.diagonal(D, dim1=-2, dim2=-1)).sum()
    return 2 * total


def mahal(decomp, y: Tensor) -> Tensor:
    """y^T J^{-1} y = |L^{-1} T y|^2.  reference mahal :461-467."""
    return sum((v * v).sum() for v in forward_sub(decomp, y))


def mahal_and_logdet(R: Tensor, O: Tensor, x: Tensor) -> Tuple[Tensor, Tensor]:
    """Fused single pass that keeps no factors.  reference mahal_and_det :380-438."""
    acc_m = x.new_zeros(())
    acc_d = x.new_zeros(())
    cur = x
    while True:
        last = R.shape[0] == 1
        if last:
            K = _chol(R)
        else:
            (m, K, F, G), (R, O) = level_step(R, O)
        acc_d = acc_d + torch.log(torch.diagonal(K, dim1=-2, dim2=-1)).sum()
        xk = _lsolve(K, cur[0::2].unsqueeze(-1)).squeeze(-1)
        acc_m = acc_m + (xk * xk).sum()
        if last:
            break
        cur = cur[1::2] - bidiag_mv(F, G, xk)
    return acc_m, 2 * acc_d


def selected_inverse(decomp) -> Tuple[Tensor, Tensor]:
    """Diagonal and lower off-diagonal blocks of J^{-1}.  reference inverse_blocks :470-503."""
    ms, Ds, Fs, Gs = decomp
    Di = torch.linalg.inv(Ds[-1])
    Sd = Di.mT @ Di
    So = Sd.new_zeros((0,) + tuple(Sd.shape[1:]))
    for k in range(len(Ds) - 2, -1, -1):
        D, F, G = Ds[k], Fs[k], Gs[k]
        Di = torch.linalg.inv(D)                                         # :484
        P = F @ Di[: F.shape[0]]                                         # F D^{-1}   :489
        Q = G @ Di[1 : 1 + G.shape[0]]                                   # G D^{-1}   :490
        mid, hi = symtri_times_bidiag(-Sd, -So, P, Q)                    # :493
        Se = Di.mT @ Di - bidiag_t_bidiag_diag(P, Q, mid, hi)            # :496
        Sd, So = interleave(Se, Sd), interleave(mid, hi.mT)              # :498-501
    return Sd, So


# --------------------------------------------------------------------------------------
# gradients of (mahal, logdet) -- closed forms (verified against autograd)
# --------------------------------------------------------------------------------------
def loglik_grads(R: Tensor, O: Tensor, x: Tensor, g_mahal: float = 1.0, g_logdet: float = 1.0):
    """Gradient of g_mahal*mahal + g_logdet*logdet wrt (R, O, x) as torch autograd through
    the reference produces it (gR symmetric):
      gR_i = g_d Sigma_ii - g_m w_i w_i^T ; gO_i = 2 g_d Sigma_{i+1,i} - 2 g_m w_{i+1} w_i^T ;
      gx = 2 g_m w,  w = J^{-1} x."""
    dec = factor(R, O)
    w = solve(dec, x)
    Sd, So = selected_inverse(dec)
    gR = g_logdet * Sd - g_mahal * torch.einsum("ia,ib->iab", w, w)
    gO = 2 * g_logdet * So - 2 * g_mahal * torch.einsum("ia,ib->iab", w[1:], w[:-1])
    return gR, gO, 2 * g_mahal * w


def loglik_grads_autograd(R: Tensor, O: Tensor, x: Tensor, g_mahal: float = 1.0, g_logdet: float = 1.0):
    """Same gradient by autograd through the fused oracle pass (what the reference's
    training step does, models.py:374-381).  This is the CPU-baseline workload."""
    R = R.detach().clone().requires_grad_(True)
    O = O.detach().clone().requires_grad_(True)
    x = x.detach().clone().requires_grad_(True)
    mh, ld = mahal_and_logdet(R, O, x)
    (g_mahal * mh + g_logdet * ld).backward()
    return (mh.detach(), ld.detach()), (R.grad, O.grad, x.grad)


# --------------------------------------------------------------------------------------
# dense helpers for small-case cross checks
# --------------------------------------------------------------------------------------
def assemble_dense(R: Tensor, O: Tensor) -> Tensor:
    n, l = R.shape[0], R.shape[1]
    J = R.new_zeros((n * l, n * l))
    for i in range(n):
        J[i * l : (i + 1) * l, i * l : (i + 1) * l] = R[i]
    for i in range(n - 1):
        J[(i + 1) * l : (i + 2) * l, i * l : (i + 1) * l] = O[i]
        J[i * l : (i + 1) * l, (i + 1) * l : (i + 2) * l] = O[i].mT
    return J


def elimination_order(n: int) -> List[int]:
    """Original indices in the order CR eliminates them (evens first, recursively);
    level k removes indices congruent to 2^k - 1 mod 2^(k+1)."""
    idx = list(range(n))
    order: List[int] = []
    while len(idx) > 1:
        order += idx[0::2]
        idx = idx[1::2]
    return order + idx


# --------------------------------------------------------------------------------------
# synthetic LEG precision blocks (inputs of the benchmark configs),
# restating models.py:152-159 (G), :199-239 (compute_PEG_precision), :254-268 (posterior)
# --------------------------------------------------------------------------------------
def leg_params(rank: int, seed: int = 0, dtype=torch.float64):
    gen = torch.Generator().manual_seed(seed)
    N = torch.eye(rank, dtype=dtype)
    A = torch.randn((rank, rank), generator=gen, dtype=dtype)
    Rm = torch.tril((A - A.T) * 0.2, diagonal=-1)
    B = torch.full((1, rank), 0.5 / rank ** 0.5, dtype=dtype)
    lam = torch.nn.functional.softplus(torch.tensor([[0.1]], dtype=dtype))
    G = N @ N.T + Rm - Rm.T + 1e-5 * torch.eye(rank, dtype=dtype)
    LLT = lam @ lam.T + 1e-9 * torch.eye(1, dtype=dtype)
    return G, B, LLT


def leg_posterior_precision(gaps: Tensor, G: Tensor, B: Tensor, LLT: Tensor):
    """(Rs, Os) of K = Sigma^{-1} + B^T (LL^T)^{-1} B for time gaps d_i (fp64)."""
    l = G.shape[0]
    eye = torch.eye(l, dtype=G.dtype)
    A = torch.matrix_exp(-0.5 * G.unsqueeze(0) * gaps.reshape(-1, 1, 1))
    At = A.mT
    left = torch.linalg.solve(eye - At @ A, At)
    right = torch.linalg.solve(eye - A @ At, A)
    Os = -right
    c1 = A @ left
    c2 = At @ right
    Rs = torch.cat([(eye + c2[0]).unsqueeze(0), eye + c1[:-1] + c2[1:], (eye + c1[-1]).unsqueeze(0)], dim=0)
    Rs = Rs + (B.T @ torch.linalg.solve(LLT, B)).unsqueeze(0)
    return Rs, Os
