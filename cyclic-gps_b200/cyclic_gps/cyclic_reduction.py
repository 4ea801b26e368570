"""Drop-in replacement for the reference module ``cyclic_gps/cyclic_reduction.py``: same
function names, argument names and return structures, computed by hand-written sm_100a
CUDA kernels (``libcrb200.so``, C ABI in ``include/crb200.h``).

Block cyclic reduction of a symmetric positive-definite block-tridiagonal matrix J with
diagonal blocks ``Rs:(n,l,l)`` and lower off-diagonal blocks ``Os:(n-1,l,l)``.

Differences from the reference, all additive:
  * every function also accepts a leading batch axis of independent series
    (``Rs:(B,n,l,l)``, ``Os:(B,n-1,l,l)``, vectors ``(B,n,l)``); scalars then have shape ``(B,)``;
  * inputs may live on the CPU or on a CUDA device; results come back on the caller's
    device (CPU inputs are copied to the current CUDA device and back -- a transfer shim,
    never a CPU compute path; without a CUDA device the calls raise);
  * ``mahal_and_det`` and ``det(decompose(...))`` carry a hand-written backward (closed-form
    gradients from a back-solve + selected inverse) instead of a torch autograd tape;
  * a non-positive-definite diagonal block goes through the reference's jitter ladder (gpytorch
    ``psd_safe_cholesky``, cyclic_reduction.py:227,306,429; see ``JITTER``) and raises
    ``NotPositiveDefiniteError`` (alias ``NotPSDError``) after the last try.

``np`` and ``torch`` are re-exported on purpose: the reference's own tests rely on
``from cyclic_gps.cyclic_reduction import *`` providing them
(tests/test_cyclic_reduction.py:250-253 of the reference).
"""
from math import ceil, floor  # noqa: F401  (names the reference module exports)
from typing import List, Tuple, Union  # noqa: F401

import numpy as np  # noqa: F401  (re-exported, see module docstring)
import torch

from . import _engine
from ._engine import NanError, NotPositiveDefiniteError, NumericalWarning  # noqa: F401

NotPSDError = NotPositiveDefiniteError   # the name the reference's callers catch (gpytorch.utils.errors.NotPSDError)

# Reference module attribute (cyclic_reduction.py:13), handed to psd_safe_cholesky as `jitter`: None = gpytorch's default
# (1e-6 in float32, 1e-8 in float64).  When a forward sweep reports a block that is not positive definite, the sweep is
# redone level by level with the reference's jitter ladder (JITTER * 10**i, i < CHOLESKY_MAX_TRIES, added to the diagonal of
# every even block of the failing level; a NumericalWarning per try), and NotPositiveDefiniteError is raised after the last
# try -- _engine.forward_sweep(jitter=...).  CHOLESKY_MAX_TRIES = 0 switches the retry off.
JITTER = None
CHOLESKY_MAX_TRIES = 3


def _forward_checked(R, O, y, **kw):
    """Forward sweep + the reference's error contract: fast path first; on a non-positive-definite report the jitter ladder."""
    pack = _engine.forward_sweep(R, O, y, **kw)
    try:
        pack.check()
        return pack
    except NotPositiveDefiniteError:
        if CHOLESKY_MAX_TRIES <= 0:
            raise
    base = JITTER if JITTER is not None else _engine.default_jitter(R.dtype)
    return _engine.forward_sweep(R, O, y, jitter=(base, CHOLESKY_MAX_TRIES), **kw)


# ---------------------------------------------------------------------------------------
# argument plumbing
# ---------------------------------------------------------------------------------------
def _check_blocks(Rs, Os):
    if not (torch.is_tensor(Rs) and torch.is_tensor(Os)):
        raise TypeError("Rs and Os must be torch tensors")
    if Rs.dim() not in (3, 4) or Os.dim() != Rs.dim():
        raise TypeError("Rs must be (n,l,l) or (B,n,l,l) and Os (n-1,l,l) or (B,n-1,l,l)")
    if Rs.shape[-1] != Rs.shape[-2] or Os.shape[-1] != Os.shape[-2] or Os.shape[-1] != Rs.shape[-1]:
        raise TypeError("blocks must be square and of one size")
    if Rs.dtype != Os.dtype:
        raise TypeError("Rs and Os must share a dtype")
    # reference: assert num_dblocks == num_offdblocks + 1  (cyclic_reduction.py:223)
    assert Rs.shape[-3] == Os.shape[-3] + 1, "need exactly one fewer off-diagonal block than diagonal blocks"
    if Rs.dim() == 4:
        assert Rs.shape[0] == Os.shape[0], "batch sizes of Rs and Os differ"


def _dev(t, dev, dtype=None):
    """Contiguous copy/view of `t` on the compute device."""
    if t is None:
        return None
    t = t.detach()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(dev, non_blocking=False).contiguous()


def _batched(t, batched):
    return t if batched else t.unsqueeze(0)


class CRDecomp(tuple):
    """``(ms, Ds, Fs, Gs)`` exactly as the reference's ``decompose`` returns it
    (cyclic_reduction.py:309), plus the device-resident packed factors and the handles
    needed for the hand-written backward."""

    def __new__(cls, ms, Ds, Fs, Gs, pack=None, Rs=None, Os=None, batched=False, caller_device=None):
        self = super().__new__(cls, (ms, Ds, Fs, Gs))
        self._pack = pack
        self._Rs, self._Os = Rs, Os
        self._batched = batched
        self._caller_device = caller_device
        return self


def _pack_of(decomp):
    """FactorPack + (batched, caller_device) for a CRDecomp or a plain 4-tuple."""
    if isinstance(decomp, CRDecomp) and decomp._pack is not None:
        return decomp._pack, decomp._batched, decomp._caller_device
    ms, Ds, Fs, Gs = decomp
    dev = _engine.require_cuda()
    caller = Ds[0].device
    batched = Ds[0].dim() == 4
    dtype, ell = Ds[0].dtype, Ds[0].shape[-1]
    msl = [int(m) for m in ms]
    B = Ds[0].shape[0] if batched else 1
    pack = _engine.FactorPack(dtype, ell, B, msl[0], msl)
    empty = torch.empty((B, 0, ell, ell), dtype=dtype, device=dev)
    for k, m in enumerate(msl):
        E, o, g = _engine.counts(m)
        pack.D.append(_batched(_dev(Ds[k], dev), batched))
        pack.F.append(_batched(_dev(Fs[k], dev), batched) if o > 0 else empty)
        pack.G.append(_batched(_dev(Gs[k], dev), batched) if g > 0 else empty)
        pack.X.append(None)
    return pack, batched, caller


def _vec_levels(pack, ycrr, batched, dev):
    out = []
    for k, m in enumerate(pack.ms):
        v = _batched(_dev(ycrr[k], dev, pack.dtype), batched)
        E = _engine.counts(m)[0]
        if v.shape[1] != E or v.shape[2] != pack.ell:
            raise ValueError(f"level {k}: expected ({E},{pack.ell}) right-hand side, got {tuple(v.shape[1:])}")
        out.append(v)
    return out


# ---------------------------------------------------------------------------------------
# block-bidiagonal helper products (reference :15-200).  Plain tensor algebra on the
# caller's device: they are tested API surface, not part of the hot path.
# U has diagonal blocks `diags` (odd node j x even node j) and super-diagonal blocks
# `offdiags` (odd node j x even node j+1); len(offdiags) == len(diags) means U is o x (o+1).
# ---------------------------------------------------------------------------------------
def UU_T(diags, offdiags):
    """Diagonal and lower off-diagonal blocks of U U^T  (reference :15-37)."""
    o, g = diags.shape[0], offdiags.shape[0]
    dd = torch.matmul(diags, diags.transpose(1, 2))
    gg = torch.matmul(offdiags, offdiags.transpose(1, 2))
    if g == o:
        dd = dd + gg
    else:
        dd = torch.cat([dd[:g] + gg, dd[g:]], dim=0)
    low = torch.matmul(diags[1:], offdiags[: max(o - 1, 0)].transpose(1, 2))
    return dd, low


def Ux(diags, offdiags, x):
    """U @ x  (reference :40-60)."""
    o, g = diags.shape[0], offdiags.shape[0]
    out = torch.matmul(diags, x[:o].unsqueeze(-1)).squeeze(-1)
    gx = torch.matmul(offdiags, x[1:g + 1].unsqueeze(-1)).squeeze(-1)
    if g == o:
        return out + gx
    return torch.cat([out[:g] + gx, out[g:]], dim=0)


def U_Tx(diags, offdiags, x):
    """U.T @ x  (reference :63-87)."""
    o, g = diags.shape[0], offdiags.shape[0]
    ft = torch.matmul(diags.transpose(1, 2), x[:o].unsqueeze(-1)).squeeze(-1)
    gt = torch.matmul(offdiags.transpose(1, 2), x[:g].unsqueeze(-1)).squeeze(-1)
    if g == o:
        return torch.cat([ft[:1], ft[1:] + gt[: o - 1], gt[o - 1:]], dim=0)
    return torch.cat([ft[:1], ft[1:] + gt], dim=0)


def SigU(sig_dblocks, sig_offdblocks, u_dblocks, u_offdblocks):
    """Diagonal and super-diagonal blocks of Sig @ U for symmetric block-tridiagonal Sig
    given by its diagonal and LOWER off-diagonal blocks  (reference :90-136)."""
    o, g = u_dblocks.shape[0], u_offdblocks.shape[0]
    mid = torch.matmul(sig_dblocks, u_dblocks)
    if o > 1:
        mid = torch.cat([mid[:1], mid[1:] + torch.matmul(sig_offdblocks, u_offdblocks[: o - 1])], dim=0)
    hi = torch.matmul(sig_dblocks[:g], u_offdblocks)
    k = min(g, o - 1)
    if k > 0:
        hi = torch.cat([hi[:k] + torch.matmul(sig_offdblocks[:k].transpose(1, 2), u_dblocks[1:k + 1]), hi[k:]], dim=0)
    return mid, hi


def UtV_diags(u_dblocks, u_offdblocks, v_dblocks, v_offdblocks):
    """Diagonal blocks of U.T @ V  (reference :139-178)."""
    o, g = u_dblocks.shape[0], u_offdblocks.shape[0]
    ff = torch.matmul(u_dblocks.transpose(1, 2), v_dblocks)
    gg = torch.matmul(u_offdblocks.transpose(1, 2), v_offdblocks)
    if g == o:
        return torch.cat([ff[:1], ff[1:] + gg[: o - 1], gg[o - 1:]], dim=0)
    return torch.cat([ff[:1], ff[1:] + gg], dim=0)


def interleave(a, b):
    """V[::2] = a, V[1::2] = b; what is left of the longer one is appended  (reference :181-200)."""
    k = min(a.shape[0], b.shape[0])
    head = torch.stack([a[:k], b[:k]], dim=1).reshape((2 * k,) + tuple(a.shape[1:]))
    return torch.cat([head, a[k:], b[k:]], dim=0)


# ---------------------------------------------------------------------------------------
# factorisation
# ---------------------------------------------------------------------------------------
def _to_caller(t, batched, caller):
    t = t if batched else t[0]
    return t if t.device == caller else t.to(caller)


def decompose_step(Rs, Os):
    """One CR level  (reference :204-259):
    returns ``(num_dblocks, Ks_even, F, G), (Rs_next, Os_next)``."""
    _check_blocks(Rs, Os)
    dev = _engine.require_cuda()
    batched, caller = Rs.dim() == 4, Rs.device
    R = _batched(_dev(Rs, dev), batched)
    O = _batched(_dev(Os, dev), batched)
    pack = _forward_checked(R, O, None, keep_factors=True, want_logdet=False, nlevels=1)
    B, ell = R.shape[0], R.shape[2]
    m = R.shape[1]
    E, o, g = _engine.counts(m)
    z = lambda rows: torch.empty((B, rows, ell, ell), dtype=R.dtype, device=dev)
    K = pack.D[0]
    F = pack.F[0] if o > 0 else z(0)
    G = pack.G[0] if g > 0 else z(0)
    Rn, On, _ = pack.rest if pack.rest is not None else (None, None, None)
    Rn = Rn if Rn is not None else z(0)
    On = On if On is not None else z(0)
    c = lambda t: _to_caller(t, batched, caller)
    return (m, c(K), c(F), c(G)), (c(Rn), c(On))


def _decompose_impl(Rs, Os):
    _check_blocks(Rs, Os)
    dev = _engine.require_cuda()
    batched, caller = Rs.dim() == 4, Rs.device
    R = _batched(_dev(Rs, dev), batched)
    O = _batched(_dev(Os, dev), batched)
    pack = _forward_checked(R, O, None, keep_factors=True)
    ell = R.shape[2]
    z = torch.empty((R.shape[0], 0, ell, ell), dtype=R.dtype, device=dev)
    if caller == dev:
        Dl, Fl, Gl = pack.D, pack.F, pack.G                 # zero-copy views of the packed device buffers
    else:
        # a caller on another device (the reference's tests pass CPU tensors): ONE copy per factor family, carved into the
        # per-level views afterwards, instead of three small copies per level
        B = R.shape[0]
        rows = [[_engine.counts(m)[i] for m in pack.ms] for i in range(3)]
        host = lambda flat, r: _engine._carve(flat.to(caller), r, (ell, ell), B) if flat is not None and flat.numel() else [z.to(caller)] * len(r)
        Dl, Fl, Gl = host(pack.D_flat, rows[0]), host(pack.F_flat, rows[1]), host(pack.G_flat, rows[2])
    c = (lambda t: t) if batched else (lambda t: t[0])
    Ds = [c(d) for d in Dl]
    Fs = [c(f) for f in Fl[:-1]]
    Gs = [c(g) if _engine.counts(pack.ms[k])[2] > 0 else c(z.to(caller)) for k, g in enumerate(Gl[:-1])]
    ms = torch.tensor(pack.ms, dtype=torch.int64)
    return CRDecomp(ms, Ds, Fs, Gs, pack=pack, Rs=Rs, Os=Os, batched=batched, caller_device=caller)


def decompose(Rs, Os):
    """Full CR factorisation  (reference :288-309): ``(ms, Ds, Fs, Gs)`` with ``ms`` an
    int64 tensor ``[n, n//2, ..., 1]`` and per-level lists of ``(E_k,l,l)``, ``(o_k,l,l)``,
    ``(g_k,l,l)`` tensors (lower-triangular Cholesky factors in ``Ds``)."""
    return _decompose_impl(Rs, Os)


# ---------------------------------------------------------------------------------------
# solves
# ---------------------------------------------------------------------------------------
def halfsolve(decomp, y):
    """L^{-1} T y in CR order: list of per-level ``(E_k, l)`` tensors  (reference :312-338)."""
    pack, batched, caller = _pack_of(decomp)
    dev = _engine.require_cuda()
    Y = _batched(_dev(y, dev, pack.dtype), batched)
    (_, X), _ = _engine.halfsolve_sweep(pack, Y)
    return [_to_caller(x, batched, y.device) for x in X]


def backhalfsolve(decomp, ycrr):
    """T^T L^{-T} y for ``ycrr`` in CR order  (reference :341-377); returns ``(n, l)``."""
    pack, batched, caller = _pack_of(decomp)
    dev = _engine.require_cuda()
    xs = _vec_levels(pack, ycrr, batched, dev)
    _, _, w = _engine.backward_sweep(pack, sigma=False, w=True, xs=xs)
    return _to_caller(w, batched, ycrr[0].device)


def _solve_on_device(pack, Y):
    (X_flat, X), _ = _engine.halfsolve_sweep(pack, Y)
    _, _, w = _engine.backward_sweep(pack, sigma=False, w=True, xs=X, xs_flat=X_flat)
    return w


class _SolveFn(torch.autograd.Function):
    """w = J^{-1} y, differentiable wrt y and wrt the (Rs, Os) that were handed to ``decompose`` (the
    reference's callers detach the solve, reference autograd through ``solve(decompose(Rs, Os), y)`` is the oracle).
    With u = J^{-1} g (one more pair of sweeps against the same factors):
        gy = u,   gR_i = -(u_i w_i^T + w_i u_i^T) / 2,   gO_i = -(u_{i+1} w_i^T + w_{i+1} u_i^T)
    (the diagonal-block gradient is symmetric, as torch's Cholesky backward makes it)."""

    @staticmethod
    def forward(ctx, Rs, Os, y, decomp):
        pack, batched, _ = _pack_of(decomp)
        dev = _engine.require_cuda()
        Y = _batched(_dev(y, dev, pack.dtype), batched)
        w = _solve_on_device(pack, Y)
        ctx.pack, ctx.batched, ctx.w = pack, batched, w
        ctx.devs = (Rs.device if Rs is not None else None, Os.device if Os is not None else None, y.device)
        return _to_caller(w, batched, y.device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        pack, batched, w = ctx.pack, ctx.batched, ctx.w
        dev = pack.device
        u = _solve_on_device(pack, _batched(_dev(g, dev, pack.dtype), batched))
        gR = gO = None
        if ctx.needs_input_grad[0]:
            uw = u.unsqueeze(-1) * w.unsqueeze(-2)
            gR = _to_caller(-0.5 * (uw + uw.transpose(-1, -2)), batched, ctx.devs[0])
        if ctx.needs_input_grad[1]:
            gO = -(u[:, 1:].unsqueeze(-1) * w[:, :-1].unsqueeze(-2) + w[:, 1:].unsqueeze(-1) * u[:, :-1].unsqueeze(-2))
            gO = _to_caller(gO, batched, ctx.devs[1])
        gy = _to_caller(u, batched, ctx.devs[2]) if ctx.needs_input_grad[2] else None
        return gR, gO, gy, None


def solve(decomp, y):
    """J^{-1} y  (reference :441-444).  Differentiable wrt ``y`` and, for a decomposition made by this module's
    ``decompose``, wrt its ``Rs`` / ``Os``."""
    pack, batched, caller = _pack_of(decomp)
    Rs = decomp._Rs if isinstance(decomp, CRDecomp) else None
    Os = decomp._Os if isinstance(decomp, CRDecomp) else None
    needs = torch.is_grad_enabled() and (y.requires_grad or any(t is not None and t.requires_grad for t in (Rs, Os)))
    if needs:
        return _SolveFn.apply(Rs, Os, y, decomp)
    dev = _engine.require_cuda()
    Y = _batched(_dev(y, dev, pack.dtype), batched)
    return _to_caller(_solve_on_device(pack, Y), batched, y.device)


def mahal(decomp, y):
    """y^T J^{-1} y  (reference :461-467)."""
    pack, batched, caller = _pack_of(decomp)
    dev = _engine.require_cuda()
    Y = _batched(_dev(y, dev, pack.dtype), batched)
    _, acc = _engine.halfsolve_sweep(pack, Y, want_mahal=True)
    out = acc.to(pack.dtype)
    out = out if batched else out[0]
    return out.to(y.device)


class _InverseBlocksFn(torch.autograd.Function):
    """Selected inverse of an existing factorisation, differentiable wrt the (Rs, Os) handed to ``decompose`` (the
    reference's autograd through ``inverse_blocks(decompose(Rs, Os))`` is the oracle).  Forward: the CUDA level kernels.
    Backward: the adjoint recursion has no kernels of its own -- the recursion is re-evaluated with differentiable torch ops
    on the device and reversed by torch autograd (``_adjoint.selected_inverse``)."""

    @staticmethod
    def forward(ctx, Rs, Os, decomp):
        pack = decomp._pack
        ctx.pack, ctx.batched, ctx.caller = pack, decomp._batched, Rs.device
        ctx.save_for_backward(Rs, Os)
        Sd, So, _ = _engine.backward_sweep(pack, sigma=True, w=False)
        return _to_caller(Sd, decomp._batched, decomp._caller_device), _to_caller(So, decomp._batched, decomp._caller_device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gSd, gSo):
        from . import _adjoint
        Rs, Os = ctx.saved_tensors
        dev, dtype = ctx.pack.device, ctx.pack.dtype
        with torch.enable_grad():
            R = _batched(_dev(Rs.detach(), dev, dtype), ctx.batched).clone().requires_grad_(True)
            O = _batched(_dev(Os.detach(), dev, dtype), ctx.batched).clone().requires_grad_(True)
            Sd, So = _adjoint.selected_inverse(R, O)
            gR, gO = torch.autograd.grad((Sd, So), (R, O), (_batched(_dev(gSd, dev, dtype), ctx.batched), _batched(_dev(gSo, dev, dtype), ctx.batched)),
                                         allow_unused=True)
        gO = gO if gO is not None else torch.zeros_like(O)
        return _to_caller(gR, ctx.batched, ctx.caller), _to_caller(gO, ctx.batched, ctx.caller), None


def inverse_blocks(decomp):
    """Diagonal and lower off-diagonal blocks of J^{-1}  (reference :470-503):
    ``(Sig_diag:(n,l,l), Sig_off:(n-1,l,l))`` with ``Sig_off[i] = (J^{-1})_{i+1,i}``.  Differentiable wrt the ``Rs`` / ``Os``
    that were handed to ``decompose`` when they require grad."""
    pack, batched, caller = _pack_of(decomp)
    if isinstance(decomp, CRDecomp) and decomp._Rs is not None and torch.is_grad_enabled() and (decomp._Rs.requires_grad or decomp._Os.requires_grad):
        return _InverseBlocksFn.apply(decomp._Rs, decomp._Os, decomp)
    Sd, So, _ = _engine.backward_sweep(pack, sigma=True, w=False)
    return _to_caller(Sd, batched, caller), _to_caller(So, batched, caller)


def solve_and_inverse_blocks(Rs, Os, y):
    """``(J^{-1} y, Sig_diag, Sig_off)`` = ``solve(dec, y)`` and ``inverse_blocks(dec)`` of ``dec = decompose(Rs, Os)`` in TWO sweeps
    instead of four: the forward sweep factorises and half-solves at once (as ``mahal_and_det`` does), one backward sweep carries
    the back-substitution and the selected inverse together.  This is the in-sample posterior of the reference
    (models.py:282-298: decompose, then solve, then inverse_blocks); not differentiable, like its caller."""
    _check_blocks(Rs, Os)
    dev = _engine.require_cuda()
    batched, caller = Rs.dim() == 4, Rs.device
    R = _batched(_dev(Rs.detach(), dev), batched)
    O = _batched(_dev(Os.detach(), dev), batched)
    Y = _batched(_dev(y.detach(), dev, R.dtype), batched)
    pack = _forward_checked(R, O, Y, keep_factors=True)
    Sd, So, w = _engine.backward_sweep(pack, sigma=True, w=True)
    c = lambda t: _to_caller(t, batched, caller)
    return c(w), c(Sd), c(So)


def check_decompose_loop_outputs(num_dblocks, Ks_even, F, G, Rs, Os):
    """Shape check of one level's outputs (reference :262-280; there only the even branch can fire -- the odd one is guarded by
    ``!= 0 & num_dblocks > 1``, which parses as a chained comparison): E = ceil(m/2) factors, floor(m/2) F blocks and reduced
    diagonal blocks, floor((m-1)/2) G blocks, floor(m/2) - 1 reduced off-diagonal blocks.  Raises AssertionError like the reference."""
    E, o, g = _engine.counts(int(num_dblocks))
    assert (Ks_even.shape[-3] == E and F.shape[-3] == o and G.shape[-3] == g and Rs.shape[-3] == o
            and Os.shape[-3] == max(o - 1, 0)), "decompose_step outputs do not match the level's block counts"


# ---------------------------------------------------------------------------------------
# log-determinant and the fused likelihood pass, with hand-written backward
# ---------------------------------------------------------------------------------------
class _LogDetFn(torch.autograd.Function):
    """log|J| of an existing factorisation; d/dR_i = Sigma_ii, d/dO_i = 2 Sigma_{i+1,i}."""

    @staticmethod
    def forward(ctx, Rs, Os, decomp):
        pack = decomp._pack
        ctx.pack, ctx.batched, ctx.caller = pack, decomp._batched, Rs.device
        out = pack.logdet.to(pack.dtype)
        out = out if decomp._batched else out[0]
        return out.to(Rs.device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        pack = ctx.pack
        dev = pack.device
        gd = g.detach().to(dev, torch.float64).reshape(-1).expand(pack.batch).contiguous()
        gm = torch.zeros_like(gd)
        gR, gO, _ = _engine.backward_sweep(pack, sigma=True, w=False, grad=(gm, gd))
        return _to_caller(gR, ctx.batched, ctx.caller), _to_caller(gO, ctx.batched, ctx.caller), None


def det(decomp):
    """log|J| = 2 sum log diag(D)  (reference :447-458; it is a log-determinant despite
    the name).  Differentiable wrt the ``Rs``/``Os`` that were handed to ``decompose``."""
    if isinstance(decomp, CRDecomp) and decomp._pack is not None:
        return _LogDetFn.apply(decomp._Rs, decomp._Os, decomp)
    ms, Ds, Fs, Gs = decomp
    total = sum(torch.log(torch.diagonal(D, dim1=-2, dim2=-1)).sum(dim=(-1, -2)) for D in Ds)
    return 2 * total


# Error contract (reference cyclic_reduction.py:429: NotPSDError is raised inside mahal_and_det): by default the
# positive-definiteness report of the forward sweep is read before mahal_and_det returns, with or without autograd.
# A training loop that is guaranteed to call backward() can opt into the deferred mode (EAGER_PD_CHECK = False): the
# report is then fetched asynchronously and raised by the backward pass after it has queued its kernels, which
# avoids draining the GPU between forward and backward (0.1-0.6 ms per step on configs[1]; bench.py opts in).
EAGER_PD_CHECK = True
# Memory: the packed factors (~3 n l^2 elements) live as long as the autograd graph, like the reference's tape, so
# backward(retain_graph=True) and repeated torch.autograd.grad calls work.  RELEASE_FACTORS_AFTER_BACKWARD = True
# frees them at the end of the first backward pass instead (a second backward then raises); bench.py opts in.
RELEASE_FACTORS_AFTER_BACKWARD = False


class _MahalAndDetFn(torch.autograd.Function):
    """(x^T J^{-1} x, log|J|) in one forward sweep; backward = back-solve + selected inverse
    with the gradient assembled inside the level-0 kernel."""

    @staticmethod
    def forward(ctx, Rs, Os, x, batched):
        dev = _engine.require_cuda()
        R = _batched(_dev(Rs, dev), batched)
        O = _batched(_dev(Os, dev), batched)
        X = _batched(_dev(x, dev, Rs.dtype), batched)
        need = any(ctx.needs_input_grad[:3])
        if need and not EAGER_PD_CHECK:
            pack = _engine.forward_sweep(R, O, X, keep_factors=need, internal=True)
            ctx.deferred = _engine.DeferredCheck([pack])         # (no jitter retry in the deferred mode)
        else:
            ctx.deferred = None
            pack = _forward_checked(R, O, X, keep_factors=need, internal=True)
        ctx.pack = pack if need else None
        ctx.batched, ctx.devs = batched, (Rs.device, Os.device, x.device)
        mh, ld = pack.mahal.to(Rs.dtype), pack.logdet.to(Rs.dtype)
        if not batched:
            mh, ld = mh[0], ld[0]
        return mh.to(Rs.device), ld.to(Rs.device)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_mahal, g_det):
        pack = ctx.pack
        if pack is None:
            raise RuntimeError("the CR factors were released by the first backward pass (RELEASE_FACTORS_AFTER_BACKWARD is set)")
        dev = pack.device
        as_vec = lambda g: g.detach().to(dev, torch.float64).reshape(-1).expand(pack.batch).contiguous()
        gR, gO, gx = _engine.backward_sweep(pack, sigma=True, w=True, grad=(as_vec(g_mahal), as_vec(g_det)))
        if RELEASE_FACTORS_AFTER_BACKWARD:
            ctx.pack = None                            # free the packed factors (~3 n l^2 elements) right away
        if ctx.deferred is not None:
            deferred, ctx.deferred = ctx.deferred, None
            deferred.wait()                            # NotPositiveDefiniteError of the forward sweep surfaces here
        b = ctx.batched
        return (_to_caller(gR, b, ctx.devs[0]), _to_caller(gO, b, ctx.devs[1]), _to_caller(gx, b, ctx.devs[2]), None)


def mahal_and_det(Rs, Os, x):
    """``(x^T J^{-1} x, log|J|)`` without keeping the factorisation  (reference :380-438).
    Under autograd the factors are kept for the hand-written backward instead of a tape."""
    _check_blocks(Rs, Os)
    batched = Rs.dim() == 4
    if x.dim() != Rs.dim() - 1 or x.shape[-2] != Rs.shape[-3] or x.shape[-1] != Rs.shape[-1]:
        raise TypeError("x must be (n,l) (or (B,n,l)) matching Rs")
    return _MahalAndDetFn.apply(Rs, Os, x, batched)
