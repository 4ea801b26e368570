"""Precision blocks of the LEG / PEG process on the device: the step in front of the CR hot path.

``peg_precision(gaps, G, shift)`` replaces ``LEGFamily.compute_PEG_precision`` / ``compute_posterior_precision`` of the
reference (cyclic_gps/models.py:181-239, 254-268): it turns the time gaps of a batch of series into the diagonal and
off-diagonal blocks ``(Rs, Os)`` with ONE kernel launch (``crb200_peg_precision_fwd``), and carries a hand-written
backward (``crb200_peg_precision_bwd``) to ``G`` and ``shift`` -- the per-gap chain through the two linear solves and
the adjoint of the matrix exponential run on the device, the host only finishes an (l x l) product.  The blocks are
written once, in the layout the level-0 CR kernel reads; with ``ts`` as the only per-row input a likelihood evaluation
ships 4 bytes per row to the GPU instead of (2 l^2 + l) elements.

G is eigendecomposed once per call on the host (l x l, as the reference's ``compute_eG`` does, model_utils.py:12-29).
Ranks up to 8 run on thread-per-gap kernels (cr_peg.cuh), ranks 9..32 on warp-per-gap kernels whose block algebra runs on
the FP64 tensor path (cr_pegw.cuh).  An ill-conditioned eigenbasis or CPU-only use fall back to the same formulas in torch ops
(``peg_precision_torch``), which is also the oracle of the tests."""
import numpy as np
import torch

from . import _engine, _native

EIG_COND_MAX = 1e7      # beyond this the eigen-form of exp(-d/2 G) loses too many digits: use torch.matrix_exp


def peg_precision_torch(gaps, G, shift=None, logdet=False):
    """Reference formulas in torch ops (differentiable by autograd): gaps (..., n-1) -> Rs (..., n, l, l), Os (..., n-1, l, l)
    [, log det of the unshifted block-tridiagonal matrix (...,), see peg_precision]."""
    eye = torch.eye(G.shape[0], dtype=G.dtype, device=G.device)
    A = torch.matrix_exp(-0.5 * G * gaps.to(G.dtype).unsqueeze(-1).unsqueeze(-1))
    At = A.transpose(-1, -2)
    fwd = torch.linalg.solve(eye - A @ At, A)            # (I - A A^T)^{-1} A
    bwd = torch.linalg.solve(eye - At @ A, At)           # (I - A^T A)^{-1} A^T
    from_prev, to_next = A @ bwd, At @ fwd
    base = eye if shift is None else eye + shift
    if gaps.shape[-1] == 0:
        diag = base.expand(gaps.shape[:-1] + (1,) + tuple(base.shape)).clone()
    else:
        diag = torch.cat([base + to_next[..., :1, :, :], base + from_prev[..., :-1, :, :] + to_next[..., 1:, :, :],
                          base + from_prev[..., -1:, :, :]], dim=-3)
    if not logdet:
        return diag, -fwd
    return diag, -fwd, -torch.linalg.slogdet(eye - A @ At)[1].sum(-1)


class _EigConsts:
    """Everything the kernels need about G, as ONE device array of doubles (+ complex V, V^{-1} on the device)."""

    def __init__(self, G, dev, G_host=None):
        Gc = (G_host if G_host is not None else G.detach().to("cpu", torch.float64)).numpy()
        h = self._host(Gc)
        l = Gc.shape[0]
        self.cond, self.nterms, self.partner, self.rows, self.folded = h["cond"], h["nterms"], h["partner"], h["rows"], h["folded"]
        # three host -> device copies in all (small problems are bound by such calls): doubles, one complex pack, one index pack
        self.buf = torch.from_numpy(h["flat"]).to(dev)
        base = self.buf.data_ptr()
        self.lam_re, self.lam_im, self.M_re, self.M_im = base, base + 8 * l, base + 16 * l, base + 16 * l + 8 * l * l * l
        self.cpack = torch.from_numpy(h["cpack"]).to(dev)
        self.V, self.Vinv, self.invdl = self.cpack[0], self.cpack[1], self.cpack[2]
        self.deg = self.invdl == 0                                  # 1 / (lam_j - lam_k) is stored as 0 exactly where the pair is degenerate
        if self.folded:
            ipack = torch.from_numpy(h["ipack"]).to(dev)
            self.re_idx, self.im_idx, self.diag = ipack[0], ipack[1], ipack[2, 0]
            self.im_sign = ipack[2, 1].to(torch.float64).view(1, l, 1, 1)

    def refresh(self, G_host):
        """New values of G into the SAME device arrays (a captured CUDA graph holds their addresses).  Returns False -- nothing
        copied -- when the structure of the spectrum (number of complex pairs / real eigenvalues, degenerate pairs handled by the
        index packs) has changed: the caller then has to build new constants and capture again."""
        h = self._host(G_host.numpy())
        if (h["nterms"], h["rows"], h["folded"], h["partner"]) != (self.nterms, self.rows, self.folded, self.partner):
            return False
        self.cond = h["cond"]
        self.buf.copy_(torch.from_numpy(h["flat"]), non_blocking=True)
        self.cpack.copy_(torch.from_numpy(h["cpack"]), non_blocking=True)
        self.deg.copy_(self.invdl == 0)
        return True

    @staticmethod
    def _host(Gc):
        # plain numpy on the host: for l x l matrices the per-call overhead of tensor ops is what this step costs
        l = Gc.shape[0]
        lam, V = np.linalg.eig(Gc)
        lam, V = lam.astype(np.complex128), V.astype(np.complex128)
        sv = np.linalg.svd(V, compute_uv=False)
        cond = float(sv[0] / sv[-1]) if sv[-1] > 0 else float("inf")
        # Order the spectrum as [eigenvalues with Im > 0 | real eigenvalues | conjugates of the first group] and make the
        # conjugate columns of V exact conjugates (G is real): the expansion of exp(cG) - I then only needs the first
        # `nterms` terms (complex ones doubled), and the kernel only accumulates the sums of one member of every pair.
        tol = 1e-12 * max(float(np.abs(lam).max()), 1e-300)
        pos = np.nonzero(lam.imag > tol)[0]
        real = np.nonzero(np.abs(lam.imag) <= tol)[0]
        npos, nreal = len(pos), len(real)
        if 2 * npos + nreal == l:
            lam = np.concatenate([lam[pos], lam[real].real.astype(np.complex128), lam[pos].conj()])
            Vr = V[:, real]
            if nreal:                                   # real eigenvalue -> real eigenvector (remove the arbitrary phase)
                ph = Vr[np.abs(Vr).argmax(axis=0), np.arange(nreal)]
                Vr = (Vr * (ph.conj() / np.abs(ph))).real.astype(np.complex128)
            V = np.concatenate([V[:, pos], Vr, V[:, pos].conj()], axis=1)
            nt = npos + nreal
            partner = list(range(nt, l)) + list(range(npos, nt)) + list(range(npos))
            weights = np.concatenate([np.full(npos, 2.0), np.ones(nreal)])
        else:                                           # (spectrum not closed under conjugation: cannot happen for a real G)
            nt, partner, weights = l, None, np.ones(l)
        Vinv = np.linalg.inv(V)
        M = (V.T[:, :, None] * Vinv[:, None, :]).reshape(l, l * l)          # M_k = V[:, k] V^{-1}[k, :]
        lam_t = np.concatenate([lam[:nt], np.zeros(l - nt, dtype=np.complex128)])
        M_t = np.concatenate([M[:nt] * weights[:, None], np.zeros((l - nt, l * l), dtype=np.complex128)])
        dl = lam[:, None] - lam[None, :]
        deg = np.abs(dl) <= 1e-9 * max(float(np.abs(lam).max()), 1e-300)
        invdl = np.where(deg, 0.0, 1.0 / np.where(deg, 1.0, dl))
        flat = np.ascontiguousarray(np.concatenate([lam_t.real, lam_t.imag, M_t.real.reshape(-1), M_t.imag.reshape(-1)]))
        # host side of the backward pass (finish_expm_adjoint): which rows of the kernel's S belong to which eigenvalue
        cplx = lam_t[:nt].imag != 0.0
        r0 = np.concatenate([[0], np.cumsum(np.where(cplx, 4, 2))])
        rows = [(int(r0[m]), int(r0[m]) + 1, int(r0[m]) + 2 if cplx[m] else -1, int(r0[m]) + 3 if cplx[m] else -1) for m in range(nt)]
        folded = partner is not None and int(r0[-1]) == 2 * l
        out = {"cond": cond, "nterms": nt, "partner": partner, "rows": rows, "folded": folded, "flat": flat,
               "cpack": np.stack([V, Vinv, invdl.astype(np.complex128)]), "ipack": None}
        if folded:
            # index form of `rows` for the device: row 2l of the (zero-extended) S stands for "no imaginary part"; eigenvalue m >= nterms
            # is the conjugate of its partner
            src = np.array(list(range(nt)) + partner[nt:], dtype=np.int64)
            base_r = r0[:-1][src]
            im0 = np.where(cplx[src], base_r + 2, 2 * l)
            im1 = np.where(cplx[src], base_r + 3, 2 * l)
            out["ipack"] = np.stack([np.stack([base_r, base_r + 1]), np.stack([im0, im1]),
                                     np.stack([np.arange(l), np.concatenate([np.ones(nt), -np.ones(l - nt)]).astype(np.int64)])]).astype(np.int64)
        return out

    def weight_rows(self, gaps, dtype):
        """E (2l, gaps): for every eigenvalue m < nterms the rows Re e^{c lam_m}, c Re e^{c lam_m} [, Im .., c Im ..], c = -gap / 2
        rounded like the kernels do (storage type first), in the row order of `rows`."""
        dev = gaps.device
        if getattr(self, "_erow", None) is None:
            part, which = np.zeros(2 * self.V.shape[0], dtype=np.int64), np.zeros(2 * self.V.shape[0], dtype=np.int64)
            for m, q in enumerate(self.rows):
                for k, r in enumerate(q):
                    if r >= 0:
                        part[r], which[r] = k, m
            self._erow = torch.from_numpy(np.stack([part, which])).to(dev)
        l, nt = self.V.shape[0], self.nterms
        c = (gaps.to(dtype) * -0.5).to(torch.float64).reshape(-1)
        lam = torch.complex(self.buf[:nt], self.buf[l:l + nt])
        ec = torch.exp(c.unsqueeze(1) * lam.unsqueeze(0))                             # (gaps, nterms)
        parts = torch.stack([ec.real, c.unsqueeze(1) * ec.real, ec.imag, c.unsqueeze(1) * ec.imag])
        return parts[self._erow[0], :, self._erow[1]]

    def finish_expm_adjoint(self, S):
        """S (2l [+1], l, l) from crb200_peg_precision_bwd -> gG (l, l) real.  T_m = V^{-1} (sum_g e^{c lam_m} gA_g^T) V for all l
        eigenvalues (conjugates by conjugation), Z_jk = (T_j - T_k)[k, j] / (lam_j - lam_k), the c-weighted sums on the
        diagonal and for equal eigenvalues, gG = Re(V^{-T} Z V^T) (Daleckii-Krein)."""
        l = self.V.shape[0]
        Sx = torch.cat([S[:2 * l], torch.zeros_like(S[:1])])
        Sc = torch.complex(Sx[self.re_idx], Sx[self.im_idx] * self.im_sign)           # [plain | c-weighted][m] (l, l)
        Tm = self.Vinv @ Sc @ self.V                                                  # (2, m, k, j)
        idx = self.diag
        Tj = Tm[0][idx, :, idx]                                                       # [j, k] = T_j[k, j]
        Tk = torch.einsum("kkj->jk", Tm[0])                                           # [j, k] = T_k[k, j]
        Tc = Tm[1][idx, :, idx]                                                       # [j, k] = T'_j[k, j]
        Z = (Tj - Tk) * self.invdl + torch.where(self.deg, Tc, torch.zeros_like(Tc))
        return (self.Vinv.transpose(0, 1) @ Z @ self.V.transpose(0, 1)).real


def builder_forward(c, gaps, shift64, dtype, want_logdet=True, out=None):
    """One launch of crb200_peg_precision_fwd: gaps (B, n-1) -> R, O, log det of the unshifted precision (B,) float64, info.
    out = (R, O): write into these contiguous (B, n, l, l) / (B, n-1, l, l) tensors (e.g. a slice of series of a larger batch)."""
    B, nm1 = gaps.shape
    n, l, dev = nm1 + 1, c.V.shape[0], gaps.device
    if out is not None:
        R, O = out
        if (tuple(R.shape) != (B, n, l, l) or tuple(O.shape) != (B, nm1, l, l) or R.dtype != dtype or O.dtype != dtype
                or R.device != dev or O.device != dev or not R.is_contiguous() or not O.is_contiguous()):
            raise ValueError("out = (R, O) must be contiguous (B, n, l, l) / (B, n-1, l, l) tensors of the dtype and device of the gaps")
    else:
        R = torch.empty((B, n, l, l), dtype=dtype, device=dev)
        O = torch.empty((B, nm1, l, l), dtype=dtype, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    ld = torch.zeros(B, dtype=torch.float64, device=dev)
    _native.peg_fwd(dtype, l, batch=B, n=n, gaps=gaps if nm1 > 0 else None, stride_gaps=gaps.stride(0) if nm1 > 0 else 0,
                    lam_re=c.lam_re, lam_im=c.lam_im, M_re=c.M_re, M_im=c.M_im, shift=shift64,
                    R=R, O=O if nm1 > 0 else None, strideR=n * l * l, strideO=nm1 * l * l, info=info, nterms=c.nterms,
                    logdet=ld if want_logdet else None)
    return R, O, ld, info


def builder_backward(c, gaps, O, gR, gO, g_logdet, dtype):
    """crb200_peg_precision_bwd + the host finish: cotangents of (R, O, logdet) -> (gG (l, l) float64, sum of the gR rows (l, l) float64)."""
    B, nm1 = gaps.shape
    n, l, dev = nm1 + 1, c.V.shape[0], gaps.device
    S = torch.zeros((2 * l + 1, l, l), dtype=torch.float64, device=dev)
    gRc = _engine._rows_contiguous(gR.to(dtype))
    if n > 1:
        wide = l > _native.peg_sum_max_ell()
        gOc = _engine._rows_contiguous(gO.to(dtype))
        gA = torch.empty((B * (n - 1), l * l), dtype=torch.float64, device=dev) if wide else None
        _native.peg_bwd(dtype, l, batch=B, n=n, gaps=gaps, stride_gaps=gaps.stride(0),
                        lam_re=c.lam_re, lam_im=c.lam_im, M_re=c.M_re, M_im=c.M_im,
                        O=O, strideO=O.stride(0), gR=gRc, gO=gOc, stride_gR=gRc.stride(0), stride_gO=gOc.stride(0), S=S,
                        nterms=c.nterms, g_logdet=g_logdet, gA=gA)
        if wide:
            # warp-per-gap kernels (ranks 9..32) return gA_g per gap: the weighted sums over the gaps are one fp64 GEMM,
            # (2l x gaps) x (gaps x l^2), with the same weights E the thread-per-gap kernel forms (include/crb200.h)
            S[:2 * l] = (c.weight_rows(gaps, dtype) @ gA).view(2 * l, l, l).transpose(1, 2)
            S[2 * l] = gRc.sum(dim=(0, 1), dtype=torch.float64)
    else:
        S[2 * l] = gRc.sum(dim=(0, 1), dtype=torch.float64)
    return c.finish_expm_adjoint(S), S[2 * l]


class _PegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gaps, G, shift, consts, want_logdet):
        B, nm1 = gaps.shape
        dtype, dev = gaps.dtype, gaps.device
        sh = shift.detach().to(dev, torch.float64).contiguous() if shift is not None else None
        R, O, ld, info = builder_forward(consts, gaps, sh, dtype, want_logdet)
        ctx.consts, ctx.dtype = consts, dtype
        ctx.save_for_backward(gaps, O)
        ctx.G_meta = (G.device, G.dtype)
        ctx.shift_meta = (shift.device, shift.dtype) if shift is not None else None
        ctx.info = info
        ctx.want_logdet = want_logdet
        if not want_logdet:
            ctx.mark_non_differentiable(ld)
        return R, O, ld

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gR, gO, gld):
        gaps, O = ctx.saved_tensors
        gG = gshift = None
        if ctx.needs_input_grad[1]:
            gl = gld.to(torch.float64).contiguous() if ctx.want_logdet else None
            gG, gsum = builder_backward(ctx.consts, gaps, O, gR, gO, gl, ctx.dtype)
            gG = gG.to(*ctx.G_meta)
            if ctx.shift_meta is not None and ctx.needs_input_grad[2]:
                gshift = gsum.to(*ctx.shift_meta)                  # the kernel sums the gR rows it stages anyway
        if gshift is None and ctx.shift_meta is not None and ctx.needs_input_grad[2]:
            gshift = gR.sum(dim=(0, 1), dtype=torch.float64).to(*ctx.shift_meta)
        return None, gG, gshift, None, None


_consts_cache = {}


def _consts_for(G, dev):
    """The eigendecomposition is redone only when G changes: the key is G's CONTENT (l x l numbers; a model rebuilds the
    tensor from its parameters on every call, so identity or version counters say nothing)."""
    G_host = G.detach().to("cpu", torch.float64)
    key = (G_host.numpy().tobytes(), tuple(G.shape), str(dev))
    hit = _consts_cache.get("last")
    if hit is not None and hit[0] == key:
        return hit[1]
    c = _EigConsts(G, dev, G_host)
    _consts_cache["last"] = (key, c)
    return c


def device_builder_available(rank: int) -> bool:
    return torch.cuda.is_available() and rank <= _native.peg_max_ell()


def peg_precision(gaps, G, shift=None, check=True, logdet=False, out=None):
    """gaps (B, n-1) or (n-1,) on a CUDA device (float32 / float64 = the dtype of the blocks), G (l,l) and shift (l,l)
    anywhere (they are tiny): returns (Rs, Os) on the device of `gaps`.  Differentiable wrt G and shift.

    ``logdet=True`` also returns log det of the UNSHIFTED block-tridiagonal matrix of every series, (B,) float64: the blocks
    are the joint precision of the Markov chain z_1 ~ N(0, I), z_{g+1} | z_g ~ N(A_g z_g, I - A_g A_g^T) for ANY A_g (P = I + B A^T
    = (I - A A^T)^{-1} is an identity), so its determinant is prod_g det(I - A_g A_g^T)^{-1}, and the kernel already holds the
    Cholesky factor of every I - A_g A_g^T.  This is the `prior_logdet` of LEGFamily.log_likelihood, for which the reference
    runs a second cyclic reduction (models.py:349-353); it is differentiable (its cotangent adds 2 g B_g to the cotangent of A_g)."""
    single = gaps.dim() == 1
    g2 = gaps.unsqueeze(0) if single else gaps
    if out is not None:
        # the blocks of a batch of series written into caller-owned storage (a pipeline that builds a large batch slice by slice
        # while the next slice is still crossing PCIe); no autograd through this form
        if single or torch.is_grad_enabled() and (G.requires_grad or (shift is not None and shift.requires_grad)):
            raise ValueError("peg_precision(out=...) takes a batch of series and is not differentiable: call it under torch.no_grad()")
        if not (g2.is_cuda and device_builder_available(G.shape[0])):
            raise ValueError("peg_precision(out=...) needs the device builder")
        consts = _consts_for(G, g2.device)
        if consts.cond > EIG_COND_MAX or not consts.folded:
            res = peg_precision_torch(g2, G.to(g2.device, g2.dtype), shift.to(g2.device, g2.dtype) if shift is not None else None, logdet=logdet)
            out[0].copy_(res[0]); out[1].copy_(res[1])
            return (out[0], out[1], res[2].to(torch.float64)) if logdet else (out[0], out[1])
        sh = shift.detach().to(g2.device, torch.float64).contiguous() if shift is not None else None
        R, O, ld, _ = builder_forward(consts, g2.contiguous(), sh, g2.dtype, logdet, out=out)
        return (R, O, ld) if logdet else (R, O)
    use_torch = not (g2.is_cuda and device_builder_available(G.shape[0]))
    if not use_torch:
        consts = _consts_for(G, g2.device)
        use_torch = consts.cond > EIG_COND_MAX or not consts.folded
    if use_torch:
        Gd = G.to(g2.device, g2.dtype)
        out = peg_precision_torch(g2, Gd, shift.to(g2.device, g2.dtype) if shift is not None else None, logdet=logdet)
        if logdet:
            out = (out[0], out[1], out[2].to(torch.float64))
    else:
        out = _PegFn.apply(g2.contiguous(), G, shift, consts, logdet)
        if not logdet:
            out = out[:2]
    return tuple(t[0] for t in out) if single else tuple(out)
