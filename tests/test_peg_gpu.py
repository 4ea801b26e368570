"""GPU: the device precision-block builder (cyclic_gps.peg, crb200_peg_precision_fwd / _bwd) against the
reference's formulas in torch ops with torch autograd (peg_precision_torch = compute_PEG_precision of the reference,
models.py:181-239, plus the posterior shift :254-268), in fp64 on the CPU."""
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu


def _model_G(l, seed, noise=1.0):
    g = torch.Generator().manual_seed(seed)
    N = torch.tril(torch.randn((l, l), generator=g, dtype=torch.float64)) * 0.4 + noise * torch.eye(l, dtype=torch.float64)
    A = torch.randn((l, l), generator=g, dtype=torch.float64)
    Rm = torch.tril((A - A.T) * 0.3, diagonal=-1)
    # (0.5 I keeps the slowest mode away from zero: 1 / gap-sized blocks are as ill-conditioned as I - A A^T, for any implementation)
    G = N @ N.T + Rm - Rm.T + (0.5 + 1e-5) * torch.eye(l, dtype=torch.float64)
    Bm = torch.randn((2, l), generator=g, dtype=torch.float64) * 0.5
    shift = Bm.T @ Bm * 3.0
    return G, shift


@pytest.mark.parametrize("dtype,tol_f,tol_g", [(torch.float64, 1e-11, 1e-9), (torch.float32, 2e-5, 1e-4)])
@pytest.mark.parametrize("l", [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16, 17, 24, 31, 32])     # <= 8: thread per gap; 9..32: warp per gap (DMMA)
def test_builder_forward_and_backward_vs_torch_autograd(l, dtype, tol_f, tol_g):
    from cyclic_gps.peg import peg_precision, peg_precision_torch
    for (B, n, seed) in ((1, 2, 1), (3, 33, 2), (2, 100, 3), (5, 31, 4), (1, 1, 5)):
        G, shift = _model_G(l, 10 * l + seed)
        gen = torch.Generator().manual_seed(seed)
        gaps = -torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64)) + 0.02
        Gr, sr = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
        R0, O0 = peg_precision_torch(gaps, Gr, sr)
        cR = torch.randn(R0.shape, generator=gen, dtype=torch.float64)
        cR = cR + cR.transpose(-1, -2) if seed % 2 else cR            # the CR backward produces symmetric gR; test both
        cO = torch.randn(O0.shape, generator=gen, dtype=torch.float64)
        ((R0 * cR).sum() + (O0 * cO).sum()).backward()
        Gd, sd = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
        R1, O1 = peg_precision(gaps.to(dtype).cuda(), Gd, sd)
        assert R1.is_cuda and R1.dtype == dtype and tuple(R1.shape) == tuple(R0.shape) and tuple(O1.shape) == tuple(O0.shape)
        assert_close(R1, R0, tol_f, f"Rs l={l} B={B} n={n}")
        if n > 1:
            assert_close(O1, O0, tol_f, f"Os l={l} B={B} n={n}")
        ((R1 * cR.to(dtype).cuda()).sum() + (O1 * cO.to(dtype).cuda()).sum()).backward()
        assert_close(sd.grad, sr.grad, tol_g, f"g shift l={l} n={n}")
        if n > 1:
            assert_close(Gd.grad, Gr.grad, tol_g, f"gG l={l} B={B} n={n}")


def test_builder_small_gaps_fp32_have_no_cancellation():
    """I - A A^T is formed from A - I (expm1), so gaps of 1e-3 keep fp32 accuracy."""
    from cyclic_gps.peg import peg_precision, peg_precision_torch
    G, shift = _model_G(8, 3)
    gaps = torch.full((2, 40), 1e-3, dtype=torch.float64)
    gaps[1] = torch.linspace(1e-3, 0.5, 40, dtype=torch.float64)
    R0, O0 = peg_precision_torch(gaps, G, shift)
    R1, O1 = peg_precision(gaps.float().cuda(), G, shift)
    assert_close(R1, R0, 2e-5, "Rs small gaps")
    assert_close(O1, O0, 2e-5, "Os small gaps")


def test_degenerate_eigenvalues_and_regular_spacing():
    """N = I, R = 0: G is a multiple of the identity (all eigenvalues equal) -- the divided differences fall back to the
    derivative form."""
    from cyclic_gps.peg import peg_precision, peg_precision_torch
    l = 4
    G = (1.0 + 1e-5) * torch.eye(l, dtype=torch.float64)
    gaps = torch.ones((2, 20), dtype=torch.float64)
    Gr = G.clone().requires_grad_(True)
    R0, O0 = peg_precision_torch(gaps, Gr, None)
    (R0.sum() + (O0 ** 2).sum()).backward()
    Gd = G.clone().requires_grad_(True)
    R1, O1 = peg_precision(gaps.cuda(), Gd, None)
    (R1.sum() + (O1 ** 2).sum()).backward()
    assert_close(R1, R0, 1e-11, "Rs")
    assert_close(Gd.grad, Gr.grad, 1e-9, "gG degenerate")


@pytest.mark.parametrize("dtype,tol_f,tol_g", [(torch.float64, 1e-11, 1e-9), (torch.float32, 2e-5, 1e-4)])
@pytest.mark.parametrize("l", [1, 2, 3, 5, 8, 10, 16, 32])
def test_builder_logdet_is_the_prior_logdet_of_the_reference(l, dtype, tol_f, tol_g):
    """logdet=True: log det of the unshifted block-tridiagonal precision as a by-product of the builder.  The
    reference obtains the same number from a second cyclic reduction, det(decompose(Sigma^{-1})) (models.py:349-353): compared
    here with the oracle's CR on the CPU, and its gradient (mixed with cotangents on Rs, Os) with torch autograd."""
    from oracle import cr_oracle as orc
    from cyclic_gps.peg import peg_precision, peg_precision_torch
    for (B, n, seed) in ((1, 2, 1), (3, 33, 2), (2, 100, 3), (4, 64, 4), (1, 1, 5)):
        G, shift = _model_G(l, 10 * l + seed)
        gen = torch.Generator().manual_seed(seed)
        gaps = -torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64)) + 0.02
        Gr, sr = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
        R0, O0, ld0 = peg_precision_torch(gaps, Gr, sr, logdet=True)
        Rp, Op = peg_precision_torch(gaps, G, None)
        for b in range(B):                                            # the identity itself: closed form == CR log-determinant
            want = orc.logdet(orc.factor(Rp[b], Op[b])) if n > 1 else torch.logdet(Rp[b, 0])
            assert abs(float(ld0[b].detach()) - float(want)) <= 1e-10 * max(1.0, abs(float(want))), (float(ld0[b].detach()), float(want))
        cR = torch.randn(R0.shape, generator=gen, dtype=torch.float64)
        cR = cR + cR.transpose(-1, -2)
        cO = torch.randn(O0.shape, generator=gen, dtype=torch.float64)
        cl = torch.randn(B, generator=gen, dtype=torch.float64)
        ((R0 * cR).sum() + (O0 * cO).sum() + (ld0 * cl).sum()).backward()
        Gd, sd = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
        R1, O1, ld1 = peg_precision(gaps.to(dtype).cuda(), Gd, sd, logdet=True)
        assert ld1.dtype == torch.float64 and tuple(ld1.shape) == (B,)
        assert_close(R1, R0, tol_f, f"Rs l={l} B={B} n={n}")
        if n > 1:
            # every pivot of chol(I - A A^T) carries one rounding error of the storage type: for large gaps the pivots are ~1 and a
            # gap's own log-determinant ~0, so the fp32 bound has an absolute part per pivot (the CR route has the same one)
            slack = 0.0 if dtype == torch.float64 else 2e-7 * l * (n - 1)
            err = float((ld1.detach().cpu() - ld0.detach()).abs().max())
            assert err <= tol_f * float(ld0.detach().abs().max()) + slack, (f"logdet l={l} B={B} n={n}", err, ld0)
        else:
            assert float(ld1.detach().abs().max()) == 0.0
        ((R1 * cR.to(dtype).cuda()).sum() + (O1 * cO.to(dtype).cuda()).sum() + (ld1 * cl.cuda()).sum()).backward()
        assert_close(sd.grad, sr.grad, tol_g, f"g shift l={l} n={n}")
        if n > 1:
            assert_close(Gd.grad, Gr.grad, tol_g, f"gG l={l} B={B} n={n}")
    # only the log-determinant is used (no cotangent reaches Rs / Os)
    G, shift = _model_G(l, 77)
    gaps = torch.rand((2, 50), dtype=torch.float64, generator=torch.Generator().manual_seed(9)) + 0.05
    Gr = G.clone().requires_grad_(True)
    peg_precision_torch(gaps, Gr, None, logdet=True)[2].sum().backward()
    Gd = G.clone().requires_grad_(True)
    peg_precision(gaps.to(dtype).cuda(), Gd, None, logdet=True)[2].sum().backward()
    assert_close(Gd.grad, Gr.grad, tol_g, "gG from logdet alone")


@pytest.mark.parametrize("l,dtype,tol_f,tol_g", [(8, torch.float32, 2e-5, 1e-4), (8, torch.float64, 1e-11, 1e-9), (5, torch.float64, 1e-11, 1e-9),
                                                (3, torch.float32, 2e-5, 1e-4), (16, torch.float64, 1e-11, 1e-9), (12, torch.float32, 2e-5, 1e-4)])
def test_builder_many_tiles_per_persistent_cta(l, dtype, tol_f, tol_g):
    """The backward kernel runs persistent CTAs whose warps reuse their shared-memory records and weight tables round after round
    (next tile prefetched under the accumulation phase): enough gaps for several rounds per CTA, ragged last tiles, against the
    same formulas in fp64 torch ops on the device."""
    from cyclic_gps.peg import peg_precision, peg_precision_torch
    B, n = (96, 2203) if l <= 8 else (24, 1201)         # 69 tiles per series, 6624 tiles: > 7 rounds of 296 CTAs x 3 warps (warp per gap: 28800 gaps)
    G, shift = _model_G(l, 5 + l)
    gen = torch.Generator(device="cuda").manual_seed(l)
    gaps = (-torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64, device="cuda")) + 0.02)
    cR = torch.randn((B, n, l, l), generator=gen, dtype=torch.float64, device="cuda")
    cR = cR + cR.transpose(-1, -2)
    cO = torch.randn((B, n - 1, l, l), generator=gen, dtype=torch.float64, device="cuda")
    cl = torch.randn(B, generator=gen, dtype=torch.float64, device="cuda")
    Gr, sr = G.cuda().requires_grad_(True), shift.cuda().requires_grad_(True)
    R0, O0, ld0 = peg_precision_torch(gaps.to(dtype).double(), Gr, sr, logdet=True)
    ((R0 * cR).sum() + (O0 * cO).sum() + (ld0 * cl).sum()).backward()
    Gd, sd = G.clone().requires_grad_(True), shift.clone().requires_grad_(True)
    R1, O1, ld1 = peg_precision(gaps.to(dtype), Gd, sd, logdet=True)
    assert_close(R1, R0, tol_f, "Rs")
    assert_close(O1, O0, tol_f, "Os")
    assert_close(ld1, ld0, tol_f, "logdet")
    ((R1 * cR.to(dtype)).sum() + (O1 * cO.to(dtype)).sum() + (ld1 * cl).sum()).backward()
    assert_close(sd.grad, sr.grad, tol_g, "g shift")
    assert_close(Gd.grad, Gr.grad, tol_g, "gG")


@pytest.mark.parametrize("l,dtype", [(8, torch.float32), (3, torch.float64), (16, torch.float64)])
def test_builder_writes_into_caller_storage_slice_by_slice(l, dtype):
    """peg_precision(out=(R, O)): a large batch built a slice of series at a time (bench.py's end-to-end pipeline builds the slice that has
    crossed PCIe while the next one is in flight) equals the one-call result bit for bit; the form is not differentiable."""
    from cyclic_gps.peg import peg_precision
    G, shift = _model_G(l, seed=11)
    B, n = 13, 141
    gaps = (torch.rand((B, n - 1), generator=torch.Generator().manual_seed(5), dtype=torch.float64) + 0.02).to(dtype).cuda()
    with torch.no_grad():
        R1, O1, ld1 = peg_precision(gaps, G, shift, logdet=True)
        R2 = torch.full((B, n, l, l), float("nan"), dtype=dtype, device="cuda")
        O2 = torch.full((B, n - 1, l, l), float("nan"), dtype=dtype, device="cuda")
        lds = []
        for lo, hi in ((0, 4), (4, 5), (5, 13)):
            r, o, ld = peg_precision(gaps[lo:hi], G, shift, logdet=True, out=(R2[lo:hi], O2[lo:hi]))
            assert r.data_ptr() == R2[lo:hi].data_ptr() and o.data_ptr() == O2[lo:hi].data_ptr()
            lds.append(ld)
    assert torch.equal(R1, R2) and torch.equal(O1, O2)
    assert_close(torch.cat(lds), ld1, 1e-12, "logdet of the slices")
    Gg = G.clone().requires_grad_(True)
    with pytest.raises(ValueError):
        peg_precision(gaps[:2], Gg, shift, out=(R2[:2], O2[:2]))                      # differentiable inputs
    with pytest.raises(ValueError):
        with torch.no_grad():
            peg_precision(gaps[:2], G, shift, out=(R2[:3], O2[:2]))                   # wrong shape
