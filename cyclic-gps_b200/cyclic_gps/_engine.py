"""Host-side driver of the CR level kernels: level plans, packed factor storage, the forward
(factor / reduce / half-solve) sweep and the backward (back-half-solve / selected inverse /
gradient assembly) sweep.  Everything here works on CUDA tensors with a leading batch axis
of independent series; ``cyclic_reduction.py`` adapts the reference's un-batched,
device-agnostic signatures onto it.

Storage (DESIGN.md "data layout"): one flat allocation per factor family holding every
level back to back; level k of D is a (B, E_k, l, l) view, F (B, o_k, l, l), G (B, g_k, l, l),
x_k (B, E_k, l).  Reduced systems ping-pong through scratch buffers that live only for the
duration of a sweep."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _native


class NotPositiveDefiniteError(RuntimeError):
    """Raised when a diagonal block met during elimination is not positive definite
    (the reference raises gpytorch's NotPSDError after its jitter ladder,
    cyclic_reduction.py:227,306,429; the jitter retry is not replicated)."""


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("the cyclic-reduction engine needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def level_sizes(n: int) -> List[int]:
    """m_0 = n, m_{k+1} = floor(m_k / 2) down to 1  (reference decompose :298-307)."""
    if n < 1:
        raise ValueError("need at least one diagonal block")
    ms = [n]
    while ms[-1] > 1:
        ms.append(ms[-1] // 2)
    return ms


def counts(m: int):
    return (m + 1) // 2, m // 2, (m - 1) // 2   # E, o, g


class FactorPack:
    """Device-resident CR factorisation of a batch of series."""

    def __init__(self, dtype, ell, batch, n, ms):
        self.dtype, self.ell, self.batch, self.n, self.ms = dtype, ell, batch, n, list(ms)
        self.D: List[torch.Tensor] = []
        self.F: List[torch.Tensor] = []
        self.G: List[torch.Tensor] = []
        self.X: List[Optional[torch.Tensor]] = []
        self.G_halo: List[torch.Tensor] = []
        self.logdet: Optional[torch.Tensor] = None     # (B,) float64: 2 * sum log diag
        self.mahal: Optional[torch.Tensor] = None      # (B,) float64
        self.info: Optional[torch.Tensor] = None
        self.rest = None                               # (R, O, y) left when the sweep stopped early
        self.halo_out = None                           # dict(Rh, yh, O) of a chunked sweep

    @property
    def nlevels(self):
        return len(self.D)

    def check(self):
        """One device->host read: raise if any diagonal block was not positive definite."""
        if self.info is None:
            return
        info = self.info.cpu()
        bad = torch.nonzero(info)
        if bad.numel():
            k = int(bad[0])
            flat = 0x7FFFFFFF - int(info[k])
            E = counts(self.ms[k])[0]
            raise NotPositiveDefiniteError(
                f"cyclic reduction: diagonal block not positive definite at level {k} "
                f"(series {flat // E}, even node {flat % E}, i.e. reduced row {2 * (flat % E)})")


def _alloc_levels(total_rows: Sequence[int], trailing, batch, dtype, device):
    """One flat buffer, carved into per-level (batch, rows_k, *trailing) views."""
    per = 1
    for t in trailing:
        per *= t
    sizes = [batch * r * per for r in total_rows]
    flat = torch.empty(sum(sizes), dtype=dtype, device=device)
    views, pos = [], 0
    for r, s in zip(total_rows, sizes):
        views.append(flat[pos:pos + s].view(batch, r, *trailing))
        pos += s
    return flat, views


def _rows_contiguous(t):
    """The kernels take an arbitrary series (batch) stride but need every series' rows back to
    back; anything else is copied."""
    if t is None or t.numel() == 0:
        return t
    inner = t[0]
    return t if inner.is_contiguous() else t.contiguous()


def forward_sweep(R: torch.Tensor, O: torch.Tensor, y: Optional[torch.Tensor], *, keep_factors: bool,
                  want_logdet: bool = True, nlevels: Optional[int] = None,
                  halo_O: Optional[torch.Tensor] = None) -> FactorPack:
    """Run CR levels 0..nlevels-1 (default: all, down to the last 1x1 system).

    R (B,n,l,l), O (B,n-1,l,l) [any strided batch axis, rows contiguous], y (B,n,l) or None."""
    B, n, ell = R.shape[0], R.shape[1], R.shape[2]
    dtype, dev = R.dtype, R.device
    R, O, y = _rows_contiguous(R), _rows_contiguous(O), _rows_contiguous(y)
    ms_all = level_sizes(n)
    L = len(ms_all) if nlevels is None else min(nlevels, len(ms_all))
    ms = ms_all[:L]
    pack = FactorPack(dtype, ell, B, n, ms)
    Es = [counts(m)[0] for m in ms]
    os_ = [counts(m)[1] for m in ms]
    gs = [counts(m)[2] for m in ms]
    if keep_factors:
        _, pack.D = _alloc_levels(Es, (ell, ell), B, dtype, dev)
        _, pack.F = _alloc_levels(os_, (ell, ell), B, dtype, dev)
        _, pack.G = _alloc_levels(gs, (ell, ell), B, dtype, dev)
    else:
        pack.D = [None] * L
        pack.F = [None] * L
        pack.G = [None] * L
    if y is not None and keep_factors:
        _, pack.X = _alloc_levels(Es, (ell,), B, dtype, dev)
    else:
        pack.X = [None] * L
    pack.logdet = torch.zeros(B, dtype=torch.float64, device=dev) if want_logdet else None
    pack.mahal = torch.zeros(B, dtype=torch.float64, device=dev) if y is not None else None
    pack.info = torch.zeros(L, dtype=torch.int32, device=dev)
    halo = halo_O is not None
    if halo:
        Rh = torch.zeros((B, ell, ell), dtype=dtype, device=dev)
        yh = torch.zeros((B, ell), dtype=dtype, device=dev)
        if keep_factors:
            pack.G_halo = list(torch.empty((L, B, ell, ell), dtype=dtype, device=dev).unbind(0))
        cur_halo = halo_O.contiguous()

    cur_R, cur_O, cur_y = R, O, y
    sR, sO = R.stride(0), (O.stride(0) if O.shape[1] > 0 else 0)
    sy = y.stride(0) if y is not None else 0
    for k, m in enumerate(ms):
        E, o, g = counts(m)
        Rn = torch.empty((B, o, ell, ell), dtype=dtype, device=dev) if o > 0 else None
        On = torch.empty((B, max(o - 1, 0), ell, ell), dtype=dtype, device=dev) if o > 1 else None
        yn = torch.empty((B, o, ell), dtype=dtype, device=dev) if (o > 0 and y is not None) else None
        fields = dict(batch=B, m=m, R=cur_R, O=cur_O if m > 1 else None, y=cur_y,
                      strideR=sR, strideO=sO, stridey=sy,
                      D=pack.D[k], F=pack.F[k] if o > 0 else None, G=pack.G[k] if g > 0 else None, xk=pack.X[k],
                      Rn=Rn, On=On, yn=yn, logdet=pack.logdet, mahal=pack.mahal, info=pack.info[k:k + 1])
        if halo:
            On_h = torch.empty((B, ell, ell), dtype=dtype, device=dev)
            fields.update(O_halo=cur_halo, G_halo=pack.G_halo[k] if keep_factors else None, On_halo=On_h, Rh_acc=Rh, yh_acc=yh)
        _native.level_fwd(dtype, ell, **fields)
        if halo:
            cur_halo = On_h
        cur_R, cur_O, cur_y = Rn, On, yn
        sR, sO, sy = o * ell * ell, max(o - 1, 0) * ell * ell, o * ell
    if pack.logdet is not None:
        pack.logdet.mul_(2.0)    # log|J| = 2 sum log diag(K)   (reference det :458, mahal_and_det :438)
    if L < len(ms_all):
        pack.rest = (cur_R, cur_O, cur_y)
    if halo:
        pack.halo_out = dict(Rh=Rh, yh=yh, O=cur_halo)
    return pack


def backward_sweep(pack: FactorPack, *, sigma: bool, w: bool, xs: Optional[Sequence[torch.Tensor]] = None,
                   grad=None, top=None, halo=None, out=None):
    """Deepest stored level first.  Returns (Sd, So, w) of level 0 (None for parts not asked).

    xs    per-level right-hand sides in CR order (default: the half-solve kept in the pack)
    grad  (gm, gd): (B,) float64 cotangents -> level 0 emits gR, gO, gx instead
    top   (Sd, So, w) of the system left by an early-stopped forward sweep
    halo  dict(Sd=(B,l,l), w=(B,l), So=(B,l,l)) values at / towards the virtual left node
    out   optional (Sd, So, w) tensors to write level 0 into"""
    B, ell, dtype = pack.batch, pack.ell, pack.dtype
    dev = pack.D[0].device
    xs = list(xs) if xs is not None else pack.X
    if w and any(x is None for x in xs):
        raise ValueError("back-solve needs the per-level right-hand sides")
    Sd_in = So_in = w_in = None
    if top is not None:
        Sd_in, So_in, w_in = top
    use_halo = halo is not None
    So_h = halo["So"] if use_halo and sigma else None
    Sd = So = wv = None
    for k in range(pack.nlevels - 1, -1, -1):
        m = pack.ms[k]
        E, o, g = counts(m)
        last = k == 0
        if last and out is not None:
            Sd, So, wv = out
        else:
            Sd = torch.empty((B, m, ell, ell), dtype=dtype, device=dev) if sigma else None
            So = torch.empty((B, max(m - 1, 0), ell, ell), dtype=dtype, device=dev) if sigma else None
            wv = torch.empty((B, m, ell), dtype=dtype, device=dev) if w else None
        fields = dict(batch=B, m=m, D=pack.D[k], F=pack.F[k] if o > 0 else None, G=pack.G[k] if g > 0 else None,
                      xk=xs[k] if w else None, Sd_in=Sd_in, So_in=So_in if o > 1 else None, w_in=w_in,
                      Sd_out=Sd, So_out=So if m > 1 else None, w_out=wv,
                      strideSd=m * ell * ell, strideSo=max(m - 1, 0) * ell * ell, stridew=m * ell,
                      gm=None, gd=None, grad_mode=0)
        if last and grad is not None:
            fields.update(gm=grad[0], gd=grad[1], grad_mode=1)
        if use_halo:
            So_h_out = torch.empty((B, ell, ell), dtype=dtype, device=dev) if sigma else None
            fields.update(G_halo=pack.G_halo[k], Sd_halo=halo["Sd"] if sigma else None, w_halo=halo["w"] if w else None,
                          So_halo_in=So_h if (sigma and o > 0) else None, So_halo_out=So_h_out)
        _native.level_bwd(dtype, ell, **fields)
        if use_halo and sigma:
            So_h = So_h_out
        Sd_in, So_in, w_in = Sd, So, wv
    if use_halo:
        return Sd, So, wv, So_h
    return Sd, So, wv


def halfsolve_sweep(pack: FactorPack, y: torch.Tensor, want_mahal: bool = False):
    """x_k for every level against stored factors (reference halfsolve :312-338)."""
    B, ell, dtype, dev = pack.batch, pack.ell, pack.dtype, y.device
    Es = [counts(m)[0] for m in pack.ms]
    _, X = _alloc_levels(Es, (ell,), B, dtype, dev)
    acc = torch.zeros(B, dtype=torch.float64, device=dev) if want_mahal else None
    cur, sy = y, y.stride(0)
    for k, m in enumerate(pack.ms):
        E, o, g = counts(m)
        yn = torch.empty((B, o, ell), dtype=dtype, device=dev) if o > 0 else None
        _native.level_halfsolve(dtype, ell, batch=B, m=m, D=pack.D[k], F=pack.F[k] if o > 0 else None,
                                G=pack.G[k] if g > 0 else None, y=cur, stridey=sy, xk=X[k], yn=yn, mahal=acc)
        cur, sy = yn, o * ell
    return X, acc
