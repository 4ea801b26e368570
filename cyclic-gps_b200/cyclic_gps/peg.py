"""Precision blocks of the LEG / PEG process on the device: the step in front of the CR hot path.

``peg_precision(gaps, G, shift)`` replaces ``LEGFamily.compute_PEG_precision`` / ``compute_posterior_precision`` of the
reference (cyclic_gps/models.py:181-239, 254-268): it turns the time gaps of a batch of series into the diagonal and
off-diagonal blocks ``(Rs, Os)`` with ONE kernel launch (``crb200_peg_precision_fwd``), and carries a hand-written
backward (``crb200_peg_precision_bwd``) to ``G`` and ``shift`` -- the per-gap chain through the two linear solves and
the adjoint of the matrix exponential run on the device, the host only finishes an (l x l) product.  The blocks are
written once, in the layout the level-0 CR kernel reads; with ``ts`` as the only per-row input a likelihood evaluation
ships 4 bytes per row to the GPU instead of (2 l^2 + l) elements.

G is eigendecomposed once per call on the host (l x l, as the reference's ``compute_eG`` does, model_utils.py:12-29).
Block sizes above ``crb200_peg_max_ell()`` (8), an ill-conditioned eigenbasis or CPU-only use fall back to the same
formulas in torch ops (``peg_precision_torch``), which is also the oracle of the tests."""
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

    def __init__(self, G, dev):
        l = G.shape[0]
        Gc = G.detach().to(torch.float64).cpu()
        lam, V = torch.linalg.eig(Gc)
        self.cond = float(torch.linalg.cond(V))
        # Order the spectrum as [eigenvalues with Im > 0 | real eigenvalues | conjugates of the first group] and make the
        # conjugate columns of V exact conjugates (G is real): the expansion of exp(cG) - I then only needs the first
        # `nterms` terms (complex ones doubled), and the kernel only produces the rows j < nterms of Z -- the others follow
        # from Z[j', k'] = conj(Z[j, k]) with ' = conjugate partner.
        tol = 1e-12 * float(lam.abs().max().clamp_min(1e-300))
        pos = [k for k in range(l) if float(lam[k].imag) > tol]
        real = [k for k in range(l) if abs(float(lam[k].imag)) <= tol]
        if len(pos) + len(real) + len(pos) == l:
            lam = torch.cat([lam[pos], lam[real].real.to(lam.dtype), lam[pos].conj()])
            Vr = V[:, real]
            if len(real):                               # real eigenvalue -> real eigenvector (remove the arbitrary phase)
                piv = Vr.abs().argmax(dim=0)
                ph = Vr[piv, torch.arange(len(real))]
                Vr = (Vr * (ph.conj() / ph.abs())).real.to(V.dtype)
            V = torch.cat([V[:, pos], Vr, V[:, pos].conj()], dim=1)
            self.nterms = len(pos) + len(real)
            self.partner = list(range(self.nterms, l)) + list(range(len(pos), self.nterms)) + list(range(len(pos)))
            weights = torch.tensor([2.0] * len(pos) + [1.0] * len(real), dtype=torch.float64)
        else:                                           # (spectrum not closed under conjugation: cannot happen for a real G)
            self.nterms, self.partner, weights = l, None, torch.ones(l, dtype=torch.float64)
        Vinv = torch.linalg.inv(V)
        M = torch.einsum("rk,kc->krc", V, Vinv).reshape(l, l * l)
        nt = self.nterms
        lam_t = torch.cat([lam[:nt], torch.zeros(l - nt, dtype=lam.dtype)])
        M_t = torch.cat([M[:nt] * weights.unsqueeze(1), torch.zeros((l - nt, l * l), dtype=M.dtype)])
        dl = lam.unsqueeze(1) - lam.unsqueeze(0)
        deg = dl.abs() <= 1e-9 * lam.abs().max().clamp_min(1e-300)
        invdl = torch.where(deg, torch.zeros_like(dl), 1.0 / torch.where(deg, torch.ones_like(dl), dl))
        parts = [lam_t.real, lam_t.imag, M_t.real, M_t.imag]
        flat = torch.cat([p.reshape(-1).to(torch.float64) for p in parts]).contiguous()
        self.buf = flat.to(dev)
        base, off, ptrs = self.buf.data_ptr(), 0, []
        for p in parts:
            ptrs.append(base + 8 * off)
            off += p.numel()
        self.lam_re, self.lam_im, self.M_re, self.M_im = ptrs
        # host side of the backward pass (finish_expm_adjoint): which rows of the kernel's S belong to which eigenvalue
        rows, r = [], 0
        for m in range(nt):
            cplx = float(lam_t[m].imag) != 0.0
            rows.append((r, r + 1, r + 2 if cplx else -1, r + 3 if cplx else -1))
            r += 4 if cplx else 2
        self.folded = self.partner is not None and r == 2 * l
        self.rows = rows
        self.V, self.Vinv = V.to(dev), Vinv.to(dev)
        self.invdl, self.deg = invdl.to(dev), deg.to(dev)
        if self.folded:
            # index form of `rows` for the device: row 2l of the (zero-extended) S stands for "no imaginary part"; eigenvalue m >= nterms
            # is the conjugate of its partner
            z = 2 * l
            ridx = [[q[0] for q in rows], [q[1] for q in rows]]
            iidx = [[q[2] if q[2] >= 0 else z for q in rows], [q[3] if q[3] >= 0 else z for q in rows]]
            src = list(range(nt)) + [self.partner[m] for m in range(nt, l)]
            self.re_idx = torch.tensor([[ridx[w][m] for m in src] for w in range(2)], device=dev)
            self.im_idx = torch.tensor([[iidx[w][m] for m in src] for w in range(2)], device=dev)
            self.im_sign = torch.tensor([1.0] * nt + [-1.0] * (l - nt), dtype=torch.float64, device=dev).view(1, l, 1, 1)
            self.diag = torch.arange(l, device=dev)

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


class _PegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gaps, G, shift, consts, want_logdet):
        B, nm1 = gaps.shape
        n, l, dtype, dev = nm1 + 1, G.shape[0], gaps.dtype, gaps.device
        R = torch.empty((B, n, l, l), dtype=dtype, device=dev)
        O = torch.empty((B, nm1, l, l), dtype=dtype, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        ld = torch.zeros(B, dtype=torch.float64, device=dev)
        sh = shift.detach().to(dev, torch.float64).contiguous() if shift is not None else None
        _native.peg_fwd(dtype, l, batch=B, n=n, gaps=gaps if nm1 > 0 else None, stride_gaps=gaps.stride(0) if nm1 > 0 else 0,
                        lam_re=consts.lam_re, lam_im=consts.lam_im, M_re=consts.M_re, M_im=consts.M_im, shift=sh,
                        R=R, O=O if nm1 > 0 else None, strideR=n * l * l, strideO=nm1 * l * l, info=info, nterms=consts.nterms,
                        logdet=ld if want_logdet else None)
        ctx.consts, ctx.meta = consts, (B, n, l, dtype)
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
        B, n, l, dtype = ctx.meta
        c = ctx.consts
        dev = gaps.device
        gG = gshift = None
        if ctx.needs_input_grad[1]:
            S = torch.zeros((2 * l + 1, l, l), dtype=torch.float64, device=dev)
            if n > 1:
                gRc = _engine._rows_contiguous(gR.to(dtype))
                gOc = _engine._rows_contiguous(gO.to(dtype))
                gl = gld.to(torch.float64).contiguous() if ctx.want_logdet else None
                _native.peg_bwd(dtype, l, batch=B, n=n, gaps=gaps, stride_gaps=gaps.stride(0),
                                lam_re=c.lam_re, lam_im=c.lam_im, M_re=c.M_re, M_im=c.M_im,
                                O=O, strideO=O.stride(0), gR=gRc, gO=gOc, stride_gR=gRc.stride(0), stride_gO=gOc.stride(0), S=S,
                                nterms=c.nterms, g_logdet=gl)
            gG = c.finish_expm_adjoint(S).to(*ctx.G_meta)
            if n > 1 and ctx.shift_meta is not None and ctx.needs_input_grad[2]:
                gshift = S[2 * l].to(*ctx.shift_meta)              # the kernel sums the gR rows it stages anyway
        if gshift is None and ctx.shift_meta is not None and ctx.needs_input_grad[2]:
            gshift = gR.sum(dim=(0, 1), dtype=torch.float64).to(*ctx.shift_meta)
        return None, gG, gshift, None, None


_consts_cache = {}


def _consts_for(G, dev):
    """The eigendecomposition is redone only when G changes: the key is G's CONTENT (l x l numbers; a model rebuilds the
    tensor from its parameters on every call, so identity or version counters say nothing)."""
    key = (G.detach().to("cpu", torch.float64).numpy().tobytes(), tuple(G.shape), str(dev))
    hit = _consts_cache.get("last")
    if hit is not None and hit[0] == key:
        return hit[1]
    c = _EigConsts(G, dev)
    _consts_cache["last"] = (key, c)
    return c


def device_builder_available(rank: int) -> bool:
    return torch.cuda.is_available() and rank <= _native.peg_max_ell()


def peg_precision(gaps, G, shift=None, check=True, logdet=False):
    """gaps (B, n-1) or (n-1,) on a CUDA device (float32 / float64 = the dtype of the blocks), G (l,l) and shift (l,l)
    anywhere (they are tiny): returns (Rs, Os) on the device of `gaps`.  Differentiable wrt G and shift.

    ``logdet=True`` also returns log det of the UNSHIFTED block-tridiagonal matrix of every series, (B,) float64: the blocks
    are the joint precision of the Markov chain z_1 ~ N(0, I), z_{g+1} | z_g ~ N(A_g z_g, I - A_g A_g^T) for ANY A_g (P = I + B A^T
    = (I - A A^T)^{-1} is an identity), so its determinant is prod_g det(I - A_g A_g^T)^{-1}, and the kernel already holds the
    Cholesky factor of every I - A_g A_g^T.  This is the `prior_logdet` of LEGFamily.log_likelihood, for which the reference
    runs a second cyclic reduction (models.py:349-353); it is differentiable (its cotangent adds 2 g B_g to the cotangent of A_g)."""
    single = gaps.dim() == 1
    g2 = gaps.unsqueeze(0) if single else gaps
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
