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


ACC_SLOTS = 32


class NotPositiveDefiniteError(RuntimeError):
    """Raised when a diagonal block met during elimination is not positive definite
    (the reference raises gpytorch's NotPSDError after its jitter ladder,
    cyclic_reduction.py:227,306,429; the jitter retry is not replicated)."""


class NanError(RuntimeError):
    """A diagonal block contains NaNs (gpytorch's NanError in the reference's psd_safe_cholesky)."""


class NumericalWarning(RuntimeWarning):
    """Jitter was added to the diagonal to make a block positive definite (gpytorch's NumericalWarning)."""


def default_jitter(dtype) -> float:
    """gpytorch.settings.cholesky_jitter defaults used by the reference's psd_safe_cholesky (cyclic_reduction.py:13, :227)."""
    return 1e-6 if dtype == torch.float32 else 1e-8


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
        # per-level lists D, F, G, X: (B, rows_k, ...) views of the packed buffers, built on first use (a sweep
        # itself only needs the packed buffers; creating ~4 L views costs more host time than a small sweep)
        self._levels = {}
        self._specs = {}
        self.G_halo: List[torch.Tensor] = []
        self.D_flat = self.F_flat = self.G_flat = self.X_flat = self.G_halo_flat = None   # packed storage behind the lists
        self.logdet: Optional[torch.Tensor] = None     # (B,) float64: 2 * sum log diag
        self.mahal: Optional[torch.Tensor] = None      # (B,) float64
        self.info: Optional[torch.Tensor] = None
        self.rest = None                               # (R, O, y) left when the sweep stopped early
        self.halo_out = None                           # dict(Rh, yh, O) of a chunked sweep

    @property
    def nlevels(self):
        return len(self.ms)

    @property
    def device(self):
        return self.D_flat.device if self.D_flat is not None else self.D[0].device

    def _get(self, name):
        if name not in self._levels:
            spec = self._specs.get(name)
            if spec is None:
                self._levels[name] = []
            else:
                flat, rows, trailing = spec
                self._levels[name] = [None] * len(rows) if flat is None else _carve(flat, rows, trailing, self.batch)
        return self._levels[name]

    D = property(lambda self: self._get("D"), lambda self, v: self._levels.__setitem__("D", v))
    F = property(lambda self: self._get("F"), lambda self, v: self._levels.__setitem__("F", v))
    G = property(lambda self: self._get("G"), lambda self, v: self._levels.__setitem__("G", v))
    X = property(lambda self: self._get("X"), lambda self, v: self._levels.__setitem__("X", v))

    def check(self, info_host: Optional[torch.Tensor] = None):
        """One device->host read: raise if any diagonal block was not positive definite.
        `info_host`: this pack's info already on the host (see `DeferredCheck`)."""
        if self.info is None:
            return
        info = self.info.cpu() if info_host is None else info_host
        bad = torch.nonzero(info)
        if bad.numel():
            k = int(bad[0])
            flat = 0x7FFFFFFF - int(info[k])
            E = counts(self.ms[k])[0]
            tail = ""
            if getattr(self, "jitter_exhausted", None):
                tail = f" after repeatedly adding jitter up to {self.jitter_exhausted:.1e}"
            raise NotPositiveDefiniteError(
                f"cyclic reduction: diagonal block not positive definite at level {k} "
                f"(series {flat // E}, even node {flat % E}, i.e. reduced row {2 * (flat % E)})" + tail)


class DeferredCheck:
    """Positive-definiteness reports of several sweeps, fetched with ONE asynchronous device->host copy
    into pinned memory.  `wait()` blocks only until that copy is done (not until the stream is idle), so a
    caller can queue further work behind the sweeps and look at the reports later without draining the GPU."""

    def __init__(self, packs: Sequence["FactorPack"]):
        self.packs = [p for p in packs if p is not None and p.info is not None]
        self.host = self.event = None
        if self.packs:
            flat = torch.cat([p.info for p in self.packs])
            self.host = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
            self.host.copy_(flat, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record()
            self._keep = flat                     # the source must outlive the copy

    def wait(self):
        if self.event is None:
            return
        self.event.synchronize()
        self.event = self._keep = None
        pos = 0
        for p in self.packs:
            k = p.info.numel()
            p.check(self.host[pos:pos + k])
            pos += k


class SideStream:
    """A second CUDA stream for work that is independent of what the current stream is doing (the other half
    of a batch of series, the ragged tail of a chunked series).  The latency-bound deep levels of one sweep
    then run under the bandwidth-bound levels of the other.  Tensors that cross streams are registered with
    the caching allocator (`record_stream`).  Inactive (everything stays on the current stream) on CPU tensors."""
    _streams = {}

    def __init__(self, dev: torch.device, active: bool = True):
        self.on = bool(active) and dev.type == "cuda"
        if self.on:
            key = dev.index if dev.index is not None else torch.cuda.current_device()
            if key not in SideStream._streams:
                SideStream._streams[key] = torch.cuda.Stream(device=dev)
            self.side = SideStream._streams[key]
            self.main = torch.cuda.current_stream(dev)

    def fork(self, *inputs):
        """The side stream may start: everything queued on the main stream so far is a dependency."""
        if self.on:
            self.side.wait_stream(self.main)
            for t in inputs:
                if t is not None:
                    t.record_stream(self.side)

    def stream(self):
        import contextlib
        return torch.cuda.stream(self.side) if self.on else contextlib.nullcontext()

    def join(self, *outputs):
        """The main stream waits for the side stream; `outputs` were allocated there and are used here next."""
        if self.on:
            self.main.wait_stream(self.side)
            for t in outputs:
                if t is not None:
                    t.record_stream(self.main)


def _carve(flat, total_rows: Sequence[int], trailing, batch):
    """Per-level (batch, rows_k, *trailing) views of a packed buffer."""
    per = 1
    for t in trailing:
        per *= t
    views, pos = [], 0
    for r in total_rows:
        size = batch * r * per
        views.append(flat[pos:pos + size].view(batch, r, *trailing))
        pos += size
    return views


def _alloc_levels(total_rows: Sequence[int], trailing, batch, dtype, device):
    """One flat buffer, carved into per-level (batch, rows_k, *trailing) views."""
    per = 1
    for t in trailing:
        per *= t
    flat = torch.empty(batch * sum(total_rows) * per, dtype=dtype, device=device)
    return flat, _carve(flat, total_rows, trailing, batch)


def _flat_only(total_rows: Sequence[int], trailing, batch, dtype, device):
    per = 1
    for t in trailing:
        per *= t
    return torch.empty(batch * sum(total_rows) * per, dtype=dtype, device=device)


class _Workspace:
    """Scratch buffers of one sweep carved out of ONE allocation and handed to the library as raw addresses
    (no per-buffer tensor objects: on small problems the host, not the GPU, is the bottleneck)."""

    def __init__(self, sizes_bytes: Sequence[int], device, zero: bool = False):
        offs, pos = [], 0
        for sz in sizes_bytes:
            offs.append(pos)
            pos += (sz + 255) & ~255
        self.buf = (torch.zeros if zero else torch.empty)(max(pos, 1), dtype=torch.uint8, device=device)
        base = self.buf.data_ptr()
        self.sizes = list(sizes_bytes)
        self.offs = offs
        self.ptrs = [base + o if sz > 0 else None for o, sz in zip(offs, sizes_bytes)]

    def view(self, i, dtype, shape):
        sz = self.sizes[i]
        return self.buf[self.offs[i]:self.offs[i] + sz].view(dtype).view(shape)


def _rows_contiguous(t):
    """The kernels take an arbitrary series (batch) stride but need every series' rows back to
    back; anything else is copied."""
    if t is None or t.numel() == 0:
        return t
    inner = t[0]
    return t if inner.is_contiguous() else t.contiguous()


def forward_sweep(R: torch.Tensor, O: torch.Tensor, y: Optional[torch.Tensor], *, keep_factors: bool,
                  want_logdet: bool = True, nlevels: Optional[int] = None,
                  halo_O: Optional[torch.Tensor] = None, jitter=None, internal: bool = False) -> FactorPack:
    """Run CR levels 0..nlevels-1 (default: all, down to the last 1x1 system).

    R (B,n,l,l), O (B,n-1,l,l) [any strided batch axis, rows contiguous], y (B,n,l) or None.
    The level loop runs inside libcrb200 (crb200_sweep_fwd); when bench.py's launch tracer is
    active the per-level entries are used instead so that every launch can be timed.

    jitter = (base, max_tries): error-recovery path that mirrors gpytorch's psd_safe_cholesky as the reference calls it
    once per level (cyclic_reduction.py:227, :306, :429): levels run one at a time; when a level reports a block that
    is not positive definite, `base * 10**i` (i = 0..max_tries-1) is added to the diagonal of ALL even blocks of that
    level (the whole batch handed to that Cholesky call) and the level is redone, with a NumericalWarning per try;
    NotPositiveDefiniteError after the last try, NanError if the blocks contain NaNs.  One host sync per level.

    internal = True: the factors go to `backward_sweep` and nowhere else (mahal_and_det under autograd, the graph runners).
    Where the kernels offer it (crb200_tri_stride: float32, ell = 8) the blocks that never leave the library -- D, the reduced
    diagonal blocks, Sigma_d of the inner levels -- are then stored as packed lower triangles (`pack.tri`): 14 % fewer bytes
    per step.  `pack.D[k]` is not a valid (B, E, l, l) view in that case."""
    B, n, ell = R.shape[0], R.shape[1], R.shape[2]
    dtype, dev = R.dtype, R.device
    bs = ell * ell
    R, O, y = _rows_contiguous(R), _rows_contiguous(O), _rows_contiguous(y)
    ms_all = level_sizes(n)
    L = len(ms_all) if nlevels is None else min(nlevels, len(ms_all))
    ms = ms_all[:L]
    pack = FactorPack(dtype, ell, B, n, ms)
    pks = _native.tri_stride(dtype, ell) if (internal and halo_O is None and jitter is None and L == len(ms_all)) else 0
    pack.tri = pks > 0
    dbs = pks if pack.tri else bs                       # elements per diagonal block of the reduced systems
    Es = [counts(m)[0] for m in ms]
    os_ = [counts(m)[1] for m in ms]
    gs = [counts(m)[2] for m in ms]
    es = torch.empty((), dtype=dtype).element_size()
    if keep_factors:
        pack.D_flat = _flat_only(Es, (ell, ell), B, dtype, dev)
        pack.F_flat = _flat_only(os_, (ell, ell), B, dtype, dev)
        pack.G_flat = _flat_only(gs, (ell, ell), B, dtype, dev)
    else:
        pack.D_flat = pack.F_flat = pack.G_flat = None
    pack._specs["D"] = (pack.D_flat, Es, (ell, ell))
    pack._specs["F"] = (pack.F_flat, os_, (ell, ell))
    pack._specs["G"] = (pack.G_flat, gs, (ell, ell))
    pack.X_flat = _flat_only(Es, (ell,), B, dtype, dev) if (y is not None and keep_factors) else None
    pack._specs["X"] = (pack.X_flat, Es, (ell,))
    # scalar accumulators: (B, ACC_SLOTS) so that the per-CTA atomics of a long series do not all hit one address;
    # they and the per-level failure words share one zero-filled allocation
    slots = ACC_SLOTS if n >= 4096 else 1
    zws = _Workspace([B * slots * 8 if want_logdet else 0, B * slots * 8 if y is not None else 0, L * 4], dev, zero=True)
    acc_ld = zws.view(0, torch.float64, (B, slots)) if want_logdet else None
    acc_mh = zws.view(1, torch.float64, (B, slots)) if y is not None else None
    pack.info = zws.view(2, torch.int32, (L,))
    halo = halo_O is not None
    Rh = yh = None
    pack.G_halo_flat = None
    if halo:
        Rh = torch.zeros((B, ell, ell), dtype=dtype, device=dev)
        yh = torch.zeros((B, ell), dtype=dtype, device=dev)
        if keep_factors:
            pack.G_halo_flat = torch.empty((L, B, ell, ell), dtype=dtype, device=dev)
            pack.G_halo = list(pack.G_halo_flat.unbind(0))
        halo_O = halo_O.contiguous()
    # ping-pong scratch for the reduced systems: slot 0 <- levels 0,2,.. ; slot 1 <- levels 1,3,..
    r0, r1 = n // 2, n // 4
    has_y = y is not None
    ws = _Workspace([B * r0 * bs * es, B * r1 * bs * es,                                   # scrR
                     B * r0 * bs * es if r0 > 1 else 0, B * r1 * bs * es if r1 > 1 else 0,  # scrO
                     B * r0 * ell * es if has_y else 0, B * r1 * ell * es if has_y else 0,  # scry
                     B * bs * es if halo else 0, B * bs * es if halo else 0], dev)          # On_halo
    pack._ws = ws                                       # keeps `rest` / `halo_out` views alive
    scrR, scrO, scry, On_h = ws.ptrs[0:2], ws.ptrs[2:4], ws.ptrs[4:6], ws.ptrs[6:8]
    sR, sO = R.stride(0), (O.stride(0) if O.shape[1] > 0 else 0)
    sy = y.stride(0) if y is not None else 0

    if jitter is not None and halo:
        raise ValueError("the jitter ladder is not available for halo (chunk-partitioned) sweeps")
    if not _native.tracing() and jitter is None:
        _native.sweep_fwd(dtype, ell, batch=B, n=n, nlevels=L, R=R, O=O if n > 1 else None, y=y,
                          strideR=sR, strideO=sO, stridey=sy,
                          D=pack.D_flat, F=pack.F_flat if (keep_factors and pack.F_flat.numel()) else None,
                          G=pack.G_flat if (keep_factors and pack.G_flat.numel()) else None, X=pack.X_flat,
                          scrR=scrR, scrO=scrO, scry=scry, logdet=acc_ld, mahal=acc_mh, acc_slots=slots, info=pack.info,
                          O_halo=halo_O, G_halo=pack.G_halo_flat, On_halo=On_h, Rh_acc=Rh, yh_acc=yh, tri=1 if pack.tri else 0)
    else:
        shp = lambda i, rows, *tr: ws.view(i, dtype, (B * rows,) + tr) if ws.sizes[i] else None
        scrR = (shp(0, r0, ell, ell), shp(1, r1, ell, ell))
        scrO = (shp(2, r0, ell, ell), shp(3, r1, ell, ell))
        scry = (shp(4, r0, ell), shp(5, r1, ell))
        On_h = (shp(6, 1, ell, ell), shp(7, 1, ell, ell))
        cur_R, cur_O, cur_y, cur_halo = R, O, y, halo_O
        for k, m in enumerate(ms):
            E, o, g = counts(m)
            slot = k & 1
            fields = dict(batch=B, m=m, R=cur_R, O=cur_O if m > 1 else None, y=cur_y,
                          strideR=sR, strideO=sO, stridey=sy,
                          D=pack.D[k], F=pack.F[k] if o > 0 else None, G=pack.G[k] if g > 0 else None, xk=pack.X[k],
                          Rn=scrR[slot] if o > 0 else None, On=scrO[slot] if o > 1 else None,
                          yn=scry[slot] if (o > 0 and y is not None) else None,
                          logdet=acc_ld, mahal=acc_mh, acc_slots=slots, info=pack.info[k:k + 1],
                          tri=((1 if k > 0 else 0) | 2) if pack.tri else 0)
            if halo:
                fields.update(O_halo=cur_halo, G_halo=pack.G_halo[k] if keep_factors else None, On_halo=On_h[slot], Rh_acc=Rh, yh_acc=yh)
                cur_halo = On_h[slot]
            snap = None
            if jitter is not None:                   # the scalars of a failed attempt must not stay in the accumulators
                snap = (acc_ld.clone() if acc_ld is not None else None, acc_mh.clone() if acc_mh is not None else None)
            _native.level_fwd(dtype, ell, **fields)
            if jitter is not None and int(pack.info[k]) != 0:
                base, tries = jitter
                Rb = cur_R if k == 0 else cur_R[:B * m].view(B, m, ell, ell)
                if bool(torch.isnan(Rb[:, 0::2]).any()):
                    raise NanError(f"cyclic reduction: NaN in a diagonal block at level {k}")
                Rj = Rb.contiguous().clone() if k == 0 else Rb.clone()
                prev, ok = 0.0, False
                for i in range(tries):
                    new = base * (10 ** i)
                    Rj[:, 0::2].diagonal(dim1=-2, dim2=-1).add_(new - prev)
                    prev = new
                    import warnings
                    warnings.warn(f"cyclic reduction level {k}: block not positive definite, added jitter of {new:.1e} to the diagonal",
                                  NumericalWarning)
                    pack.info[k:k + 1].zero_()
                    if snap[0] is not None:
                        acc_ld.copy_(snap[0])
                    if snap[1] is not None:
                        acc_mh.copy_(snap[1])
                    fields.update(R=Rj, strideR=m * bs)
                    _native.level_fwd(dtype, ell, **fields)
                    if int(pack.info[k]) == 0:
                        ok = True
                        break
                if not ok:
                    pack.jitter_exhausted = prev
                    pack.check()
            cur_R, cur_O, cur_y = fields["Rn"], fields["On"], fields["yn"]
            sR, sO, sy = o * dbs, max(o - 1, 0) * bs, o * ell
    if slots == 1:
        pack.mahal = acc_mh.view(B) if acc_mh is not None else None
        pack.logdet = acc_ld.view(B).mul_(2.0) if acc_ld is not None else None
    else:
        pack.mahal = acc_mh.sum(dim=1) if acc_mh is not None else None
        pack.logdet = acc_ld.sum(dim=1).mul_(2.0) if acc_ld is not None else None   # log|J| = 2 sum log diag(K)  (reference :458, :438)
    last = (L - 1) & 1
    if L < len(ms_all):
        o = ms[-1] // 2
        pack.rest = (ws.view(last, dtype, (B * (r1 if last else r0), ell, ell))[:B * o].view(B, o, ell, ell),
                     ws.view(2 + last, dtype, (B * (r1 if last else r0), ell, ell))[:B * (o - 1)].view(B, o - 1, ell, ell) if o > 1 else None,
                     ws.view(4 + last, dtype, (B * (r1 if last else r0), ell))[:B * o].view(B, o, ell) if y is not None else None)
    if halo:
        pack.halo_out = dict(Rh=Rh, yh=yh, O=ws.view(6 + last, dtype, (B, ell, ell)))
    return pack


def _flat_levels(levels: Sequence[torch.Tensor]):
    """Pack per-level (B, rows_k, ...) tensors back to back (the layout crb200_sweep_bwd reads)."""
    return torch.cat([t.reshape(-1) for t in levels]) if len(levels) else None


def backward_sweep(pack: FactorPack, *, sigma: bool, w: bool, xs: Optional[Sequence[torch.Tensor]] = None,
                   grad=None, top=None, halo=None, out=None, xs_flat: Optional[torch.Tensor] = None):
    """Deepest stored level first.  Returns (Sd, So, w) of level 0 (None for parts not asked).

    xs    per-level right-hand sides in CR order (default: the half-solve kept in the pack)
    grad  (gm, gd): (B,) float64 cotangents -> level 0 emits gR, gO, gx instead
    top   (Sd, So, w) of the system left by an early-stopped forward sweep
    halo  dict(Sd=(B,l,l), w=(B,l), So=(B,l,l)) values at / towards the virtual left node
    out   optional (Sd, So, w) tensors to write level 0 into: (B,m,l,l), (B,m-1,l,l), (B,m,l), any batch
          stride, rows contiguous"""
    B, ell, dtype = pack.batch, pack.ell, pack.dtype
    dev = pack.device
    bs = ell * ell
    L = pack.nlevels
    n = pack.ms[0]
    if w:
        if xs is None:
            if pack.X_flat is None:
                raise ValueError("back-solve needs the per-level right-hand sides")
            X_flat, X_levels = pack.X_flat, pack.X
        else:
            X_levels = [x.contiguous() for x in xs]
            X_flat = xs_flat if xs_flat is not None else _flat_levels(X_levels)   # xs_flat: the levels already packed
    else:
        X_flat, X_levels = None, [None] * L
    D_flat = getattr(pack, "D_flat", None)
    if D_flat is None:
        pack.D_flat, pack.F_flat, pack.G_flat = _flat_levels(pack.D), _flat_levels(pack.F), _flat_levels(pack.G)
    top_Sd = top_So = top_w = None
    if top is not None:
        top_Sd, top_So, top_w = (t.contiguous() if t is not None else None for t in top)
    if out is not None:
        Sd, So, wv = out
    else:
        Sd = torch.empty((B, n, ell, ell), dtype=dtype, device=dev) if sigma else None
        So = torch.empty((B, max(n - 1, 0), ell, ell), dtype=dtype, device=dev) if sigma else None
        wv = torch.empty((B, n, ell), dtype=dtype, device=dev) if w else None
    m1 = pack.ms[1] if L > 1 else 0
    m2 = pack.ms[2] if L > 2 else 0
    es = torch.empty((), dtype=dtype).element_size()
    use_halo = halo is not None
    tri = bool(getattr(pack, "tri", False))
    if tri and (use_halo or top is not None or not sigma):
        raise ValueError("packed factors (forward_sweep(internal=True)) only feed the full backward sweep")
    dbs = _native.tri_stride(dtype, ell) if tri else bs
    if tri and dbs == 0:
        raise RuntimeError("the factors were packed but CRB200_TRI / the kernel variant changed since")
    hs = B * bs * es if (use_halo and sigma) else 0
    # ping-pong scratch of the descending (Sigma_d, Sigma_o, w), slot [1] <- odd levels, slot [0] <- even levels >= 2
    ws = _Workspace([B * m2 * bs * es if sigma else 0, B * m1 * bs * es if sigma else 0,
                     B * m2 * bs * es if sigma else 0, B * m1 * bs * es if sigma else 0,
                     B * m2 * ell * es if w else 0, B * m1 * ell * es if w else 0, hs, hs], dev)
    scrSd, scrSo, scrw, So_h = ws.ptrs[0:2], ws.ptrs[2:4], ws.ptrs[4:6], ws.ptrs[6:8]
    So_h_out = torch.empty((B, ell, ell), dtype=dtype, device=dev) if (use_halo and sigma) else None
    gm, gd = (grad if grad is not None else (None, None))
    stride0 = lambda t, default: (t.stride(0) if (t is not None and t.dim() > 1 and t.shape[0] > 1) else default)

    if not _native.tracing():
        _native.sweep_bwd(dtype, ell, batch=B, n=n, nlevels=L,
                          D=pack.D_flat, F=pack.F_flat if (pack.F_flat is not None and pack.F_flat.numel()) else None,
                          G=pack.G_flat if (pack.G_flat is not None and pack.G_flat.numel()) else None, X=X_flat,
                          top_Sd=top_Sd if sigma else None, top_So=top_So if sigma else None, top_w=top_w if w else None,
                          Sd_out=Sd, So_out=So if (So is not None and n > 1) else None, w_out=wv,
                          strideSd=stride0(Sd, n * bs), strideSo=stride0(So, max(n - 1, 0) * bs), stridew=stride0(wv, n * ell),
                          scrSd=scrSd, scrSo=scrSo, scrw=scrw, gm=gm, gd=gd, grad_mode=1 if grad is not None else 0,
                          G_halo=pack.G_halo_flat if use_halo else None, Sd_halo=halo["Sd"].contiguous() if (use_halo and sigma) else None,
                          w_halo=halo["w"].contiguous() if (use_halo and w) else None,
                          So_halo_in=halo["So"].contiguous() if (use_halo and sigma) else None, So_halo=So_h, So_halo_out=So_h_out,
                          tri=1 if tri else 0)
    else:
        shp = lambda i, rows, *tr: ws.view(i, dtype, (B * rows,) + tr) if ws.sizes[i] else None
        scrSd = (shp(0, m2, ell, ell), shp(1, m1, ell, ell))
        scrSo = (shp(2, m2, ell, ell), shp(3, m1, ell, ell))
        scrw = (shp(4, m2, ell), shp(5, m1, ell))
        So_h = (shp(6, 1, ell, ell), shp(7, 1, ell, ell))
        Sd_in, So_in, w_in = top_Sd, top_So, top_w
        so_h = halo["So"].contiguous() if (use_halo and sigma) else None
        for k in range(L - 1, -1, -1):
            m = pack.ms[k]
            E, o, g = counts(m)
            slot = k & 1
            if k == 0:
                oSd, oSo, ow = Sd, So, wv
                sSd, sSo, sw = stride0(Sd, n * bs), stride0(So, max(n - 1, 0) * bs), stride0(wv, n * ell)
            else:
                oSd, oSo, ow = scrSd[slot], scrSo[slot], scrw[slot]
                sSd, sSo, sw = m * dbs, max(m - 1, 0) * bs, m * ell
            fields = dict(batch=B, m=m, tri=(1 | (2 if k > 0 else 0)) if tri else 0, D=pack.D[k], F=pack.F[k] if o > 0 else None, G=pack.G[k] if g > 0 else None,
                          xk=X_levels[k] if w else None, Sd_in=Sd_in if sigma else None, So_in=So_in if (sigma and o > 1) else None,
                          w_in=w_in if w else None, Sd_out=oSd, So_out=oSo if m > 1 else None, w_out=ow,
                          strideSd=sSd, strideSo=sSo, stridew=sw, gm=None, gd=None, grad_mode=0)
            if k == 0 and grad is not None:
                fields.update(gm=gm, gd=gd, grad_mode=1)
            if use_halo:
                h_out = (So_h_out if k == 0 else So_h[slot]) if sigma else None
                fields.update(G_halo=pack.G_halo[k], Sd_halo=halo["Sd"].contiguous() if sigma else None,
                              w_halo=halo["w"].contiguous() if w else None,
                              So_halo_in=so_h if (sigma and o > 0) else None, So_halo_out=h_out)
                so_h = h_out
            _native.level_bwd(dtype, ell, **fields)
            Sd_in, So_in, w_in = oSd, oSo, ow
    if use_halo:
        return Sd, So, wv, So_h_out
    return Sd, So, wv


def halfsolve_sweep(pack: FactorPack, y: torch.Tensor, want_mahal: bool = False):
    """x_k for every level against stored factors (reference halfsolve :312-338).  Returns (X levels, mahal)."""
    B, ell, dtype, dev = pack.batch, pack.ell, pack.dtype, y.device
    n = pack.ms[0]
    y = _rows_contiguous(y)
    Es = [counts(m)[0] for m in pack.ms]
    X_flat, X = _alloc_levels(Es, (ell,), B, dtype, dev)
    acc = torch.zeros(B, dtype=torch.float64, device=dev) if want_mahal else None
    if getattr(pack, "D_flat", None) is None:
        pack.D_flat, pack.F_flat, pack.G_flat = _flat_levels(pack.D), _flat_levels(pack.F), _flat_levels(pack.G)
    r0, r1 = n // 2, n // 4
    scry = (torch.empty((B * r0, ell), dtype=dtype, device=dev) if r0 else None,
            torch.empty((B * r1, ell), dtype=dtype, device=dev) if r1 else None)
    if not _native.tracing():
        _native.sweep_halfsolve(dtype, ell, batch=B, n=n, nlevels=pack.nlevels, D=pack.D_flat,
                                F=pack.F_flat if (pack.F_flat is not None and pack.F_flat.numel()) else None,
                                G=pack.G_flat if (pack.G_flat is not None and pack.G_flat.numel()) else None,
                                y=y, stridey=y.stride(0), X=X_flat, scry=scry, mahal=acc)
    else:
        cur, sy = y, y.stride(0)
        for k, m in enumerate(pack.ms):
            E, o, g = counts(m)
            yn = scry[k & 1] if o > 0 else None
            _native.level_halfsolve(dtype, ell, batch=B, m=m, D=pack.D[k], F=pack.F[k] if o > 0 else None,
                                    G=pack.G[k] if g > 0 else None, y=cur, stridey=sy, xk=X[k], yn=yn, mahal=acc)
            cur, sy = yn, o * ell
    pack_x = (X_flat, X)
    return pack_x, acc
