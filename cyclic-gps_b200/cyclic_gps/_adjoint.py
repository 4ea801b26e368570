"""Backward pass of ``inverse_blocks``.

The selected inverse is produced by the CUDA level kernels (``_engine.backward_sweep``).  Its ADJOINT -- cotangents on the
blocks of J^{-1} back to (Rs, Os) -- is not a selected-inverse-shaped problem (g J = -J^{-1} C J^{-1} restricted to the
tridiagonal pattern needs the reverse recursions of both sweeps with arbitrary cotangents on D, F, G), no caller in the
reference differentiates it (models.py:282-298 detaches), and so it has no hand-written kernels: the backward pass
re-evaluates the recursion below with differentiable torch ops ON THE DEVICE (batched Cholesky / triangular solves /
matmuls, one group of ops per level, log2 n levels) and lets torch autograd reverse it.  The recursion is the one the
kernels implement (reference cyclic_reduction.py:204-259 forward, :470-503 selected inverse): eliminate the even nodes,
recurse on the odd ones, then
    Sigma_{2e+1,2e} = -(S_d[e] P_e + S_o[e-1] Q_{e-1}),   Sigma_{2e-1,2e} = -(S_o[e-1]^T P_e + S_d[e-1] Q_{e-1}),
    Sigma_{2e,2e}   = K_e^{-T} K_e^{-1} - P_e^T Sigma_{2e+1,2e} - Q_{e-1}^T Sigma_{2e-1,2e},
with P_e = F_e K_e^{-1}, Q_{e-1} = G_{e-1} K_e^{-1} and (S_d, S_o) the selected inverse of the reduced system."""
import torch


def _interleave(a, b):
    """a (B,p,...), b (B,q,...) with q in {p-1, p}: a0 b0 a1 b1 ..."""
    p, q = a.shape[1], b.shape[1]
    out = a.new_empty((a.shape[0], p + q) + tuple(a.shape[2:]))
    out[:, 0::2] = a
    out[:, 1::2] = b
    return out


def selected_inverse(R, O):
    """R (B,m,l,l), O (B,m-1,l,l) -> (Sigma_d (B,m,l,l), Sigma_o (B,m-1,l,l)), Sigma_o[i] = (J^{-1})_{i+1,i}; torch ops only."""
    B, m, l, _ = R.shape
    eye = torch.eye(l, dtype=R.dtype, device=R.device)
    K = torch.linalg.cholesky(R[:, 0::2])
    Ki = torch.linalg.solve_triangular(K, eye.expand_as(K), upper=False)
    KiTKi = Ki.mT @ Ki
    if m == 1:
        return KiTKi, O
    E, o, g = (m + 1) // 2, m // 2, (m - 1) // 2
    F = O[:, 0::2] @ Ki[:, :o].mT                    # F_e = O_{2e} K_e^{-T}
    P = F @ Ki[:, :o]
    Gm = O[:, 1::2].mT @ Ki[:, 1:1 + g].mT           # G_{e-1} = O_{2e-1}^T K_e^{-T}
    Q = Gm @ Ki[:, 1:1 + g]
    Rn = R[:, 1::2] - F @ F.mT
    if g > 0:
        Rn = torch.cat([Rn[:, :g] - Gm @ Gm.mT, Rn[:, g:]], dim=1)
    On = -(F[:, 1:] @ Gm[:, :o - 1].mT)
    Sd, So = selected_inverse(Rn, On)                # o nodes
    z = lambda k: R.new_zeros((B, k, l, l))
    # N1_e = Sigma_{2e+1,2e}, e < o
    N1 = Sd @ P
    if o > 1:
        N1 = N1 + torch.cat([z(1), So @ Q[:, :o - 1]], dim=1)
    N1 = -N1
    # N2_{e-1} = Sigma_{2e-1,2e}, e = 1..g  (P_e exists for e < o)
    N2 = Sd[:, :g] @ Q
    if g > 0:
        k = min(g, o - 1)
        if k > 0:
            N2 = N2 + torch.cat([So[:, :k].mT @ P[:, 1:1 + k], z(g - k)], dim=1)
    N2 = -N2
    See = KiTKi - torch.cat([P.mT @ N1, z(E - o)], dim=1)
    if g > 0:
        See = See - torch.cat([z(1), Q.mT @ N2, z(E - 1 - g)], dim=1)
    return _interleave(See, Sd), _interleave(N1, N2.mT)
