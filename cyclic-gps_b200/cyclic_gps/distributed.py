"""Chunk-partitioned cyclic reduction of ONE long series across the GPUs of a box
(exact-order variant).

The series is cut into aligned sub-chunks of ``sub = 2**K`` rows; every rank owns a
contiguous run of sub-chunks.  During CR levels 0..K-1 a sub-chunk touches the rest of the
series only through the last node of the previous sub-chunk, which survives all K levels:
the level kernels treat it as a *left halo* (a virtual surviving node -1, see
``crb200_fwd_args.O_halo``).  After K local levels every full sub-chunk has shrunk to its last
node; those nodes, plus the halo updates and couplings produced locally, form the boundary
block-tridiagonal system (one node per sub-chunk).  It is exchanged with ONE all-gather
(NCCL over NVLink; ~ (3 l^2 + 2 l + 2) numbers per sub-chunk), reduced redundantly on every
rank, and its solution / selected inverse seed the local descent for the backward pass.
A ragged tail (n not a multiple of ``sub``) is a sub-chunk that disappears completely.

Row ownership convention: rank r holds rows [lo, hi) as
  ``R_loc (n_loc, l, l)``, ``x_loc (n_loc, l)`` and ``Oprev_loc (n_loc, l, l)`` with
  ``Oprev_loc[j] = J_{lo+j, lo+j-1}`` -- i.e. every row brings the block that couples it to
  the PREVIOUS row, so the coupling to the previous rank is local (global row 0's entry is
  ignored).

The engine (level sweeps) is injectable so that the host logic can be exercised on CPU with
gloo and an oracle-backed engine in the tests; the product default is the CUDA engine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import _engine


@dataclass
class ChunkPlan:
    n: int                 # global rows
    world: int
    sub: int               # rows per sub-chunk, a power of two
    bounds: List[Tuple[int, int]]   # per rank: [first sub-chunk, one past last sub-chunk)

    @property
    def levels(self) -> int:
        return int(math.log2(self.sub))

    @property
    def nsub(self) -> int:
        return (self.n + self.sub - 1) // self.sub

    def rows(self, rank: int) -> Tuple[int, int]:
        a, b = self.bounds[rank]
        return min(a * self.sub, self.n), min(b * self.sub, self.n)   # ranks beyond the last sub-chunk own nothing

    def full_subchunks(self, rank: int) -> int:
        lo, hi = self.rows(rank)
        return (hi - lo) // self.sub

    def tail_rows(self, rank: int) -> int:
        lo, hi = self.rows(rank)
        return (hi - lo) % self.sub

    @property
    def nboundary(self) -> int:
        return self.n // self.sub


def make_plan(n: int, world: int, sub: Optional[int] = None, per_rank: int = 12) -> ChunkPlan:
    """Aligned sub-chunks, ~`per_rank` of them per rank so that the ragged tail costs < 1/per_rank."""
    if sub is None:
        target = max(1, n // max(1, per_rank * world))
        sub = 1 << max(1, int(math.floor(math.log2(max(target, 2)))))
        sub = min(sub, 1 << 22)
    if sub < 2 or sub & (sub - 1):
        raise ValueError("sub must be a power of two >= 2")
    if n < 1 or world < 1:
        raise ValueError("need n >= 1 rows and world >= 1 ranks")
    nsub = (n + sub - 1) // sub          # may be < world: the surplus ranks get an empty row range
    base, extra = divmod(nsub, world)
    bounds, pos = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        bounds.append((pos, pos + cnt))
        pos += cnt
    return ChunkPlan(n=n, world=world, sub=sub, bounds=bounds)


def _gather(t: torch.Tensor, group, world: int) -> torch.Tensor:
    if world == 1:
        return t.unsqueeze(0)
    import torch.distributed as dist
    dev = t.device
    if t.is_cuda and dist.get_backend(group) == "gloo":    # test rigs: several ranks sharing one GPU over gloo
        t = t.cpu()
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)      # concatenation along dim 0 (NCCL and gloo agree on this form)
    return out.view((world,) + tuple(t.shape)).to(dev)


# Same switches as cyclic_reduction.EAGER_PD_CHECK / RELEASE_FACTORS_AFTER_BACKWARD, for the chunked path: by default the
# non-positive-definite report is read before the forward call returns and the factors live as long as the graph.
DEFERRED_PD_CHECK = False
RELEASE_FACTORS_AFTER_BACKWARD = False


class _Ctx:
    pass


def _Side(dev, active):
    """The ragged tail of a rank is a sweep of its own (a batch of one, ~log2(rows) latency-bound levels deep).
    It is independent of the sweep over the full sub-chunks, so on a CUDA device it runs on a second stream
    and its latency chain hides behind the bandwidth-bound levels of the main sweep (0.37 ms per step on the
    rank that owns the tail)."""
    return _engine.SideStream(dev, active)


def _finish_check(ctx, keep):
    """Non-positive-definite reports of the (up to three) sweeps of a forward pass.  By default they are looked at
    at once (one device->host read).  With DEFERRED_PD_CHECK and a backward pass to follow (`keep`), the read is
    queued asynchronously and looked at by `chunked_backward` after it has queued its own work: a blocking read
    here drains the GPU and leaves it idle while the host prepares the backward pass (~0.3 ms per step)."""
    packs = (ctx.pack, ctx.pack_tail, getattr(ctx, "bpack", None))
    engine = ctx.engine
    Deferred = getattr(engine, "DeferredCheck", None)
    if keep and DEFERRED_PD_CHECK and Deferred is not None and ctx.shape[2].type == "cuda":
        ctx.deferred = Deferred(packs)
    else:
        ctx.deferred = None
        for p in packs:
            if p is not None:
                p.check()


def chunked_forward(R_loc, Oprev_loc, x_loc, plan: ChunkPlan, rank: int, *, group=None, keep=False, engine=_engine):
    """Returns (mahal, logdet) as 0-dim float64 tensors (identical on every rank) and a context
    for `chunked_backward` when keep=True."""
    ell, dtype, dev = R_loc.shape[-1], R_loc.dtype, R_loc.device
    bs = ell * ell
    sub, K = plan.sub, plan.levels
    S = plan.full_subchunks(rank)
    tail = plan.tail_rows(rank)
    lo, hi = plan.rows(rank)
    assert R_loc.shape[0] == hi - lo == Oprev_loc.shape[0] == x_loc.shape[0]
    R_loc, Oprev_loc, x_loc = R_loc.contiguous(), Oprev_loc.contiguous(), x_loc.contiguous()
    ctx = _Ctx()
    ctx.plan, ctx.rank, ctx.group, ctx.engine, ctx.S, ctx.tail = plan, rank, group, engine, S, tail
    ctx.shape = (R_loc.shape, dtype, dev)

    width = 3 * bs + 2 * ell + 2
    S_max = max(plan.full_subchunks(r) for r in range(plan.world))

    def halo_blocks(first_row, count, stride):
        h = Oprev_loc[first_row:first_row + count * stride:stride].clone() if count > 0 else Oprev_loc[:0].clone()
        if lo == 0 and first_row == 0 and count > 0:
            h[0].zero_()                     # global row 0 has no predecessor
        return h

    # what a rank contributes to the boundary system, one row per full sub-chunk (+ one for the ragged tail):
    # [ R_last | y_last | halo dR | halo dy | coupling O | logdet | mahal ] in float64; built with three kernels
    # per sweep (cat, cast, cat) -- this runs between the latency-bound ends of two sweeps
    def boundary_rows(count, parts, pack):
        lead = torch.cat([p.reshape(count, -1) for p in parts], dim=1).double()
        return torch.cat([lead, pack.logdet.view(count, 1), pack.mahal.view(count, 1)], dim=1)

    ctx.pack = ctx.pack_tail = None
    pieces = []
    side = _Side(dev, tail > 0 and S > 0 and engine is _engine)
    side.fork(R_loc, Oprev_loc, x_loc)
    if tail > 0:
        t0 = S * sub
        Rt = R_loc[t0:].unsqueeze(0)
        xt = x_loc[t0:].unsqueeze(0)
        Ot = Oprev_loc[t0 + 1:].unsqueeze(0)
        with side.stream():
            pack_t = engine.forward_sweep(Rt, Ot, xt, keep_factors=keep, halo_O=halo_blocks(t0, 1, 1))
            h = pack_t.halo_out
            zb, zv = h["Rh"].new_zeros((1, bs)), h["Rh"].new_zeros((1, ell))
            piece_t = boundary_rows(1, (zb, zv, h["Rh"], h["yh"], zb), pack_t)
        ctx.pack_tail = pack_t
    else:
        piece_t = None
    if S > 0:
        Rb = R_loc[:S * sub].view(S, sub, ell, ell)
        xb = x_loc[:S * sub].view(S, sub, ell)
        Ob = torch.as_strided(Oprev_loc, (S, sub - 1, ell, ell), (sub * bs, bs, ell, 1), Oprev_loc.storage_offset() + bs)
        pack = engine.forward_sweep(Rb, Ob, xb, keep_factors=keep, nlevels=K, halo_O=halo_blocks(0, S, sub))
        Rl, _, yl = pack.rest
        h = pack.halo_out
        pieces.append(boundary_rows(S, (Rl, yl, h["Rh"], h["yh"], h["O"]), pack))
        ctx.pack = pack
    if S < S_max:
        pieces.append(torch.zeros((S_max - S, width), dtype=torch.float64, device=dev))
    side.join(piece_t, *(ctx.pack_tail.info, ctx.pack_tail.logdet, ctx.pack_tail.mahal) if ctx.pack_tail is not None else ())
    pieces.append(piece_t if piece_t is not None else torch.zeros((1, width), dtype=torch.float64, device=dev))
    send = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=0)       # (S_max + 1, width)

    allv = _gather(send, group, plan.world)                       # (world, S_max+1, width)
    # boundary system in global sub-chunk order
    rows = []
    tail_row = None
    for r in range(plan.world):
        rows.append(allv[r, :plan.full_subchunks(r)])
        if plan.tail_rows(r) > 0:
            tail_row = allv[r, S_max]
    full = torch.cat(rows, dim=0)                                  # (Sg, width)
    Sg = full.shape[0]
    part_ld = full[:, -2].sum() + (tail_row[-2] if tail_row is not None else 0.0)
    part_mh = full[:, -1].sum() + (tail_row[-1] if tail_row is not None else 0.0)
    ctx.Sg = Sg
    if Sg == 0:
        ctx.bpack = None
        _finish_check(ctx, keep)
        return part_mh, part_ld, ctx
    Rbnd = full[:, 0:bs].clone()
    ybnd = full[:, bs:bs + ell].clone()
    dR, dy, cpl = full[:, bs + ell:2 * bs + ell], full[:, 2 * bs + ell:2 * bs + 2 * ell], full[:, 2 * bs + 2 * ell:3 * bs + 2 * ell]
    Rbnd[:-1] += dR[1:]                                            # halo updates land on the PREVIOUS boundary node
    ybnd[:-1] += dy[1:]
    if tail_row is not None:
        Rbnd[-1] += tail_row[bs + ell:2 * bs + ell]
        ybnd[-1] += tail_row[2 * bs + ell:2 * bs + 2 * ell]
    Obnd = cpl[1:]                                                 # coupling (g-1, g) was produced by sub-chunk g
    Rbnd = Rbnd.to(dtype).contiguous().view(1, Sg, ell, ell)
    ybnd = ybnd.to(dtype).contiguous().view(1, Sg, ell)
    Obnd = Obnd.to(dtype).contiguous().view(1, Sg - 1, ell, ell)    # (a column slice of `full` is NOT row-contiguous)
    bpack = engine.forward_sweep(Rbnd, Obnd, ybnd, keep_factors=keep)
    ctx.bpack = bpack
    _finish_check(ctx, keep)
    return part_mh + bpack.mahal[0], part_ld + bpack.logdet[0], ctx


def chunked_backward(ctx, g_mahal, g_logdet):
    """Gradient of g_mahal*mahal + g_logdet*logdet wrt (R_loc, Oprev_loc, x_loc) of this rank.
    The cotangents may be Python floats or 0-dim tensors (kept on the device: no host read)."""
    plan, rank, engine, S, tail = ctx.plan, ctx.rank, ctx.engine, ctx.S, ctx.tail
    (rshape, dtype, dev) = ctx.shape
    ell = rshape[-1]
    sub = plan.sub
    n_loc = rshape[0]
    bs = ell * ell
    gR = torch.empty(rshape, dtype=dtype, device=dev)
    gO = torch.empty(rshape, dtype=dtype, device=dev)
    gx = torch.empty((n_loc, ell), dtype=dtype, device=dev)
    # boundary solution: Sigma and w at every boundary node (no gradient scaling here)
    if ctx.bpack is not None:
        Sbd, Sbo, wb = engine.backward_sweep(ctx.bpack, sigma=True, w=True)
        Sbd, Sbo, wb = Sbd[0], Sbo[0], wb[0]
    else:
        # the whole series is shorter than one sub-chunk: no boundary node exists, the tail's left halo is empty
        Sbd = torch.zeros((1, ell, ell), dtype=dtype, device=dev)
        Sbo = torch.zeros((0, ell, ell), dtype=dtype, device=dev)
        wb = torch.zeros((1, ell), dtype=dtype, device=dev)
    g0 = plan.bounds[rank][0]                                       # global index of the first local sub-chunk
    zero_b = torch.zeros((1, ell, ell), dtype=dtype, device=dev)
    zero_v = torch.zeros((1, ell), dtype=dtype, device=dev)

    def halo_for(first_g, count):
        """Sigma / w of the boundary node left of sub-chunks first_g.. and Sigma_off towards it."""
        idx = torch.arange(first_g - 1, first_g - 1 + count, device=dev)
        ok = idx >= 0
        idc = idx.clamp(min=0)
        Sd = torch.where(ok.view(-1, 1, 1), Sbd[idc], zero_b)
        w = torch.where(ok.view(-1, 1), wb[idc], zero_v)
        return Sd.contiguous(), w.contiguous(), idc, ok

    def cot(B):
        vec = lambda g: (g.detach().to(dev, torch.float64).reshape(1).expand(B).contiguous() if torch.is_tensor(g)
                         else torch.full((B,), float(g), dtype=torch.float64, device=dev))
        return vec(g_mahal), vec(g_logdet)

    side = _Side(dev, tail > 0 and S > 0 and engine is _engine)
    if tail > 0:                                                    # queued first, on the side stream (see _Side)
        t0 = S * sub
        Sd_h, w_h, idc, ok = halo_for(g0 + S, 1)
        halo = dict(Sd=Sd_h, w=w_h, So=zero_b.clone())
        out = (gR[t0:].unsqueeze(0), gO[t0 + 1:].unsqueeze(0), gx[t0:].unsqueeze(0))
        cot_t = cot(1)
        side.fork(gR, gO, gx, Sd_h, w_h, halo["So"], *cot_t)
        with side.stream():
            _, _, _, So_left_t = engine.backward_sweep(ctx.pack_tail, sigma=True, w=True, grad=cot_t, halo=halo, out=out)
            gO[t0] = So_left_t[0]
    if S > 0:
        Sd_h, w_h, idc, ok = halo_for(g0, S)
        gidx = torch.arange(g0, g0 + S, device=dev)
        # Sigma_{g, g-1} between consecutive boundary nodes seeds the halo off-diagonal descent
        So_h = torch.where(ok.view(-1, 1, 1), Sbo[idc.clamp(max=max(Sbo.shape[0] - 1, 0))] if Sbo.shape[0] > 0 else zero_b.expand(S, -1, -1), zero_b)
        top = (Sbd[gidx].unsqueeze(1).contiguous(), None, wb[gidx].unsqueeze(1).contiguous())
        halo = dict(Sd=Sd_h, w=w_h, So=So_h.contiguous())
        # level 0 writes straight into the gradient tensors (rows of a sub-chunk are contiguous; the
        # off-diagonal gradient skips the first block of every sub-chunk, which is the halo coupling)
        out = (gR[:S * sub].view(S, sub, ell, ell),
               torch.as_strided(gO, (S, sub - 1, ell, ell), (sub * bs, bs, ell, 1), gO.storage_offset() + bs),
               gx[:S * sub].view(S, sub, ell))
        _, _, _, So_left = engine.backward_sweep(ctx.pack, sigma=True, w=True, grad=cot(S), top=top, halo=halo, out=out)
        gO[0:S * sub:sub] = So_left
    side.join()
    if plan.rows(rank)[0] == 0 and n_loc > 0:
        gO[0].zero_()                                               # row 0 has no predecessor
    if getattr(ctx, "deferred", None) is not None:
        ctx.deferred.wait()                                         # raises NotPositiveDefiniteError of the forward pass
        ctx.deferred = None
    return gR, gO, gx


class ChunkedMahalAndDet(torch.autograd.Function):
    """(x^T J^{-1} x, log|J|) of one long series whose rows are spread over the ranks of `group`."""

    @staticmethod
    def forward(ctx, R_loc, Oprev_loc, x_loc, plan, rank, group, engine):
        need = any(ctx.needs_input_grad[:3])
        mh, ld, c = chunked_forward(R_loc.detach(), Oprev_loc.detach(), x_loc.detach(), plan, rank, group=group, keep=need,
                                    engine=engine or _engine)
        ctx.c = c if need else None
        return mh.to(R_loc.dtype), ld.to(R_loc.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_mahal, g_det):
        if ctx.c is None:
            raise RuntimeError("the CR factors were released by the first backward pass (RELEASE_FACTORS_AFTER_BACKWARD is set)")
        gR, gO, gx = chunked_backward(ctx.c, g_mahal, g_det)
        if RELEASE_FACTORS_AFTER_BACKWARD:
            ctx.c = None                               # release ~3 n l^2 elements of factors now, not at graph teardown
        return gR, gO, gx, None, None, None, None


def chunked_mahal_and_det(R_loc, Oprev_loc, x_loc, plan: ChunkPlan, rank: int, group=None, engine=None):
    return ChunkedMahalAndDet.apply(R_loc, Oprev_loc, x_loc, plan, rank, group, engine)
