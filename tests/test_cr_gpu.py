"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped Python API)
against the CPU oracle and the golden vectors of the unmodified reference.

Tolerances are BASELINE.json's: relative 1e-10 in fp64 and 1e-4 in fp32, where relative means
max-abs-error / max-abs-reference per tensor."""
import numpy as np
import pytest
import torch

from helpers import TOL, assert_close, relerr, split_levels
from oracle import cr_oracle as orc

pytestmark = pytest.mark.gpu


def cr():
    from cyclic_gps import cyclic_reduction
    return cyclic_reduction


def leg_inputs(l, n, dtype, seed=0, spacing="irregular"):
    g = torch.Generator().manual_seed(seed)
    G, B, LLT = orc.leg_params(l, seed=seed)
    if n == 1:
        R = (torch.eye(l, dtype=torch.float64) + B.T @ torch.linalg.solve(LLT, B)).unsqueeze(0)
        O = torch.zeros((0, l, l), dtype=torch.float64)
    else:
        gaps = (-torch.log(torch.rand(n - 1, generator=g, dtype=torch.float64)) + 0.01) if spacing == "irregular" \
            else torch.ones(n - 1, dtype=torch.float64)
        R, O = orc.leg_posterior_precision(gaps, G, B, LLT)
    x = torch.randn((n, l), generator=g, dtype=torch.float64)
    return R.to(dtype), O.to(dtype), x.to(dtype)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ------------------------------------------------------------------ one level
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("l", [1, 2, 3, 4, 5, 8, 11, 16])
def test_decompose_step_vs_oracle(l, dtype):
    c = cr()
    tol = TOL[dtype]
    for m in [2, 3, 4, 5, 6, 7, 8, 9, 31, 32, 33, 63, 64, 65, 66, 127, 130, 257]:
        if l >= 11 and m > 70:
            continue
        R, O, _ = leg_inputs(l, m, dtype, seed=m)
        (mm, K, F, G), (Rn, On) = c.decompose_step(R, O)
        (m2, K2, F2, G2), (Rn2, On2) = orc.level_step(R.double(), O.double())
        assert mm == m2 and K.device == R.device
        for a, b, name in ((K, K2, "K"), (F, F2, "F"), (G, G2, "G"), (Rn, Rn2, "Rn"), (On, On2, "On")):
            assert tuple(a.shape) == tuple(b.shape), (name, m, a.shape, b.shape)
            assert_close(a, b, tol, f"{name} l={l} m={m}")
        assert torch.equal(torch.triu(K, diagonal=1), torch.zeros_like(K))   # exact zeros above the diagonal


# ------------------------------------------------------------------ golden vectors
def _check_golden_case(g, p, tol):
    c = cr()
    R, O, x = _t(g[p + "R"]), _t(g[p + "O"]), _t(g[p + "x"])
    l = R.shape[1]
    dec = c.decompose(R, O)
    ms, Ds, Fs, Gs = dec
    assert ms.dtype == torch.int64 and np.array_equal(ms.numpy(), g[p + "ms"])
    assert len(Ds) == len(ms) and len(Fs) == len(ms) - 1 and len(Gs) == len(ms) - 1
    if (p + "D") in g.files:
        for name, lv in (("D", Ds), ("F", Fs), ("G", Gs)):
            want = split_levels(g[p + name], g[p + name + "_counts"], (l, l))
            for k, (a, b) in enumerate(zip(lv, want)):
                assert tuple(a.shape) == tuple(b.shape), (name, k)
                assert_close(a, b, tol, f"{p}{name}[{k}]")
    hs = c.halfsolve(dec, x)
    want = split_levels(g[p + "halfsolve"], g[p + "halfsolve_counts"], (l,))
    assert len(hs) == len(want)
    for a, b in zip(hs, want):
        assert_close(a, b, tol, p + "halfsolve")
    assert_close(c.solve(dec, x), g[p + "solve"], tol, p + "solve")
    assert_close(c.det(dec), g[p + "logdet"], tol, p + "det")
    assert_close(c.mahal(dec, x), g[p + "mahal"], tol, p + "mahal")
    mm, dd = c.mahal_and_det(R, O, x)
    assert mm.dim() == 0 and dd.dim() == 0 and mm.dtype == R.dtype
    assert_close(mm, g[p + "mahal_fused"], tol, p + "mahal_and_det[0]")
    assert_close(dd, g[p + "logdet_fused"], tol, p + "mahal_and_det[1]")
    Sd, So = c.inverse_blocks(dec)
    assert_close(Sd, g[p + "Sd"], tol, p + "inverse diag")
    assert_close(So, g[p + "So"], tol, p + "inverse off")
    counts = [(int(m) + 1) // 2 for m in g[p + "ms"]]
    ycrr = split_levels(g[p + "ycrr"], counts, (l,))
    assert_close(c.backhalfsolve(dec, ycrr), g[p + "backhalfsolve"], tol, p + "backhalfsolve")
    for tag, (gm, gd) in (("11", (1.0, 1.0)), ("ab", (0.3, -0.7))):
        Rr, Or, xr = [t.clone().requires_grad_(True) for t in (R, O, x)]
        m2, d2 = c.mahal_and_det(Rs=Rr, Os=Or, x=xr)
        (gm * m2 + gd * d2).backward()
        assert_close(Rr.grad, g[p + "gR_" + tag], tol, p + "gR " + tag)
        if O.numel():
            assert_close(Or.grad, g[p + "gO_" + tag], tol, p + "gO " + tag)
        assert_close(xr.grad, g[p + "gx_" + tag], tol, p + "gx " + tag)


def _golden_cases(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    return [str(c) for c in np.load(path)["cases"]]


@pytest.mark.parametrize("case", _golden_cases("leg"))
def test_golden_leg(golden, case):
    _check_golden_case(golden["leg"], case, 1e-4 if "float32" in case else 1e-10)


def test_golden_random_llt(golden):
    # the reference test's own generator (J = L L^T, ill-conditioned): same checks the reference
    # makes, with np.allclose defaults against dense numpy, plus tight comparison with the reference's CR
    g = golden["random_llt"]
    c = cr()
    for p in g["cases"]:
        p = str(p)
        _check_golden_case(g, p, 1e-8)
        R, O, x = _t(g[p + "R"]), _t(g[p + "O"]), _t(g[p + "x"])
        dec = c.decompose(R, O)
        assert np.allclose(c.det(dec).numpy(), g[p + "dense_logdet"])
        assert np.allclose(c.solve(dec, x).numpy(), g[p + "dense_solve"])


def test_golden_known_matrices(golden):
    # reference tests/test_cyclic_reduction.py:243-291 (fp32, closed forms, np.allclose defaults)
    g = golden["known"]
    c = cr()
    x = _t(g["x"])
    for name, l in (("bab", 1), ("schur", 2)):
        R, O = _t(g[name + "_R"]), _t(g[name + "_O"])
        assert R.dtype == torch.float32
        dec = c.decompose(R, O)
        mm, dd = c.mahal_and_det(R, O, x=x.reshape(-1, l))
        assert np.allclose(g[name + "_logdet_closed"], c.det(dec).numpy())
        assert np.allclose(g[name + "_logdet_closed"], dd.numpy())
        Sd, So = c.inverse_blocks(dec)
        assert np.allclose(g[name + "_inv_R_closed"], Sd.numpy())
        assert np.allclose(g[name + "_inv_O_closed"], So.numpy())
        assert np.allclose(g[name + "_mahal_closed"], mm.numpy())


# ------------------------------------------------------------------ bigger cases vs the oracle
@pytest.mark.parametrize("l,n,dtype", [(3, 1000, torch.float64), (8, 4097, torch.float32), (4, 10000, torch.float32),
                                         (16, 502, torch.float64), (2, 3000, torch.float64), (32, 70, torch.float64),
                                         (24, 40, torch.float32), (8, 1023, torch.float64)])
def test_full_path_vs_oracle(l, n, dtype):
    c = cr()
    tol = TOL[dtype]
    R, O, x = leg_inputs(l, n, dtype, seed=l * 7 + n)
    Rd, Od, xd = R.double(), O.double(), x.double()
    dec_o = orc.factor(Rd, Od)
    dec = c.decompose(R.cuda(), O.cuda())
    assert dec[1][0].is_cuda
    assert_close(c.det(dec), orc.logdet(dec_o), tol, "logdet")
    assert_close(c.solve(dec, x.cuda()), orc.solve(dec_o, xd), tol, "solve")
    assert_close(c.mahal(dec, x.cuda()), orc.mahal(dec_o, xd), tol, "mahal")
    Sd, So = c.inverse_blocks(dec)
    Sd_o, So_o = orc.selected_inverse(dec_o)
    assert_close(Sd, Sd_o, tol, "Sigma diag")
    assert_close(So, So_o, tol, "Sigma off")
    Rr, Or, xr = [t.cuda().requires_grad_(True) for t in (R, O, x)]
    mm, dd = c.mahal_and_det(Rr, Or, xr)
    (0.5 * mm - 1.5 * dd).backward()
    gR, gO, gx = orc.loglik_grads(Rd, Od, xd, 0.5, -1.5)
    assert_close(mm, orc.mahal(dec_o, xd), tol, "mahal fused")
    assert_close(Rr.grad, gR, tol, "gR")
    assert_close(Or.grad, gO, tol, "gO")
    assert_close(xr.grad, gx, tol, "gx")


@pytest.mark.parametrize("l,n,dtype,variant", [
    (8, 2050, torch.float32, 1), (8, 2050, torch.float32, 2), (8, 2050, torch.float32, 3),      # lane-per-row / thread-per-node / column-split
    (8, 777, torch.float64, 1), (8, 777, torch.float64, 3),
    (4, 1500, torch.float64, 1), (4, 1500, torch.float64, 2), (4, 1500, torch.float64, 3),
    (4, 3001, torch.float32, 1), (4, 3001, torch.float32, 2),
    (3, 500, torch.float64, 1), (3, 500, torch.float64, 2), (5, 300, torch.float32, 1), (5, 300, torch.float32, 2),
    (9, 700, torch.float32, 1), (9, 700, torch.float32, 2), (10, 700, torch.float32, 1), (10, 700, torch.float32, 2),   # upper end of thread-per-node
    (6, 700, torch.float64, 1), (6, 700, torch.float64, 2), (7, 700, torch.float64, 1), (7, 700, torch.float64, 2),
    # warp-per-node DMMA family (variant 4): every padded size 8 / 16 / 24 / 32, odd sizes, both storage types;
    # lane-per-row (variant 1) kept as the cross-check at the same sizes
    (8, 700, torch.float32, 4), (8, 700, torch.float64, 4), (9, 300, torch.float64, 4), (11, 333, torch.float32, 4),
    (12, 500, torch.float64, 4), (13, 301, torch.float64, 4), (16, 1100, torch.float64, 4), (16, 1100, torch.float32, 4),
    (17, 260, torch.float32, 4), (21, 200, torch.float64, 4), (24, 300, torch.float64, 4), (24, 300, torch.float32, 4),
    (27, 130, torch.float32, 4), (31, 150, torch.float64, 4), (32, 260, torch.float64, 4), (32, 260, torch.float32, 4),
    (16, 300, torch.float64, 1), (24, 100, torch.float64, 1)])
def test_every_kernel_family_vs_oracle(l, n, dtype, variant):
    """The same contract is implemented by up to four kernel families (include/crb200.h, `variant`);
    force each one and compare the whole path (forward, backward, selected inverse, solve) with the oracle."""
    from cyclic_gps import _native
    c = cr()
    tol = TOL[dtype]
    R, O, x = leg_inputs(l, n, dtype, seed=3 * l + n + variant)
    Rd, Od, xd = R.double(), O.double(), x.double()
    dec_o = orc.factor(Rd, Od)
    _native.VARIANT = variant
    try:
        Rr, Or, xr = [t.cuda().requires_grad_(True) for t in (R, O, x)]
        mm, dd = c.mahal_and_det(Rr, Or, xr)
        (1.25 * mm + 0.75 * dd).backward()
        dec = c.decompose(R.cuda(), O.cuda())
        w = c.solve(dec, x.cuda())
        Sd, So = c.inverse_blocks(dec)
    finally:
        _native.VARIANT = 0
    gR, gO, gx = orc.loglik_grads(Rd, Od, xd, 1.25, 0.75)
    Sd_o, So_o = orc.selected_inverse(dec_o)
    for got, want, name in ((mm, orc.mahal(dec_o, xd), "mahal"), (dd, orc.logdet(dec_o), "logdet"), (w, orc.solve(dec_o, xd), "solve"),
                            (Sd, Sd_o, "Sigma diag"), (So, So_o, "Sigma off"), (Rr.grad, gR, "gR"), (Or.grad, gO, "gO"), (xr.grad, gx, "gx")):
        assert_close(got, want, tol, f"{name} (variant {variant})")
    for k, (a, b) in enumerate(zip(dec[1], dec_o[1])):
        assert_close(a, b, tol, f"D[{k}] (variant {variant})")


def test_per_level_entries_equal_the_sweep_entries():
    """libcrb200 offers the level loop inside the library (crb200_sweep_*) and one entry per level
    (crb200_level_*, used when bench.py times every launch).  Both must give the same numbers."""
    from cyclic_gps import _native

    class Tracer:                      # any object with begin/end switches the engine to the per-level entries
        enabled = True
        launches = 0

        def begin(self, kind, dtype, ell, batch, m):
            Tracer.launches += 1
            return None

        def end(self, tok):
            return None

    c = cr()
    for (l, n, dtype) in ((8, 1500, torch.float32), (3, 700, torch.float64), (16, 130, torch.float64)):
        R, O, x = (t.cuda() for t in leg_inputs(l, n, dtype, seed=n))
        results = []
        for traced in (False, True):
            _native.TRACE = Tracer() if traced else None
            try:
                Rr, Or, xr = [t.clone().requires_grad_(True) for t in (R, O, x)]
                mm, dd = c.mahal_and_det(Rr, Or, xr)
                (mm - 2 * dd).backward()
                dec = c.decompose(R, O)
                results.append((mm, dd, Rr.grad, Or.grad, xr.grad, c.solve(dec, x), c.mahal(dec, x), *c.inverse_blocks(dec), *c.halfsolve(dec, x)))
            finally:
                _native.TRACE = None
        for a, b in zip(*results):
            assert_close(a, b, 1e-12 if dtype == torch.float64 else 1e-5, "per-level vs sweep")
    assert Tracer.launches > 0


def test_det_of_decompose_is_differentiable():
    c = cr()
    R, O, _ = leg_inputs(4, 300, torch.float64, seed=5)
    Rr, Or = R.clone().requires_grad_(True), O.clone().requires_grad_(True)
    ld = c.det(c.decompose(Rr, Or))
    (2.0 * ld).backward()
    gR, gO, _ = orc.loglik_grads(R, O, torch.zeros(300, 4, dtype=torch.float64), 0.0, 2.0)
    assert_close(Rr.grad, gR, 1e-10, "d logdet / dR")
    assert_close(Or.grad, gO, 1e-10, "d logdet / dO")


def test_batched_equals_loop():
    c = cr()
    B, n, l = 5, 77, 3
    cases = [leg_inputs(l, n, torch.float64, seed=40 + b) for b in range(B)]
    R = torch.stack([k[0] for k in cases]).cuda()
    O = torch.stack([k[1] for k in cases]).cuda()
    x = torch.stack([k[2] for k in cases]).cuda()
    Rr, Or, xr = R.clone().requires_grad_(True), O.clone().requires_grad_(True), x.clone().requires_grad_(True)
    mm, dd = c.mahal_and_det(Rr, Or, xr)
    assert mm.shape == (B,) and dd.shape == (B,)
    wts = torch.linspace(0.5, 1.5, B, dtype=torch.float64, device="cuda")
    ((wts * mm).sum() + (dd / wts).sum()).backward()
    dec = c.decompose(R, O)
    w = c.solve(dec, x)
    Sd, So = c.inverse_blocks(dec)
    for b in range(B):
        Rb, Ob, xb = cases[b]
        dec_o = orc.factor(Rb, Ob)
        assert_close(mm[b], orc.mahal(dec_o, xb), 1e-10)
        assert_close(dd[b], orc.logdet(dec_o), 1e-10)
        assert_close(w[b], orc.solve(dec_o, xb), 1e-10)
        sd, so = orc.selected_inverse(dec_o)
        assert_close(Sd[b], sd, 1e-10)
        assert_close(So[b], so, 1e-10)
        gR, gO, gx = orc.loglik_grads(Rb, Ob, xb, float(wts[b]), float(1 / wts[b]))
        assert_close(Rr.grad[b], gR, 1e-10)
        assert_close(Or.grad[b], gO, 1e-10)
        assert_close(xr.grad[b], gx, 1e-10)


def test_edge_cases():
    c = cr()
    # n = 1: ms = [1], no F / G   (must work)
    R, O, x = leg_inputs(3, 1, torch.float64)
    dec = c.decompose(R, O)
    ms, Ds, Fs, Gs = dec
    assert ms.tolist() == [1] and len(Ds) == 1 and Fs == [] and Gs == []
    assert_close(c.det(dec), torch.logdet(R[0]), 1e-12)
    assert_close(c.solve(dec, x), torch.linalg.solve(R[0], x[0]).unsqueeze(0), 1e-12)
    assert_close(c.inverse_blocks(dec)[0], torch.linalg.inv(R[0]).unsqueeze(0), 1e-12)
    assert c.inverse_blocks(dec)[1].shape == (0, 3, 3)
    # n = 2, l = 1
    R = torch.tensor([[[4.0]], [[9.0]]], dtype=torch.float64)
    O = torch.tensor([[[1.0]]], dtype=torch.float64)
    mm, dd = c.mahal_and_det(R, O, torch.tensor([[1.0], [2.0]], dtype=torch.float64))
    J = torch.tensor([[4.0, 1.0], [1.0, 9.0]], dtype=torch.float64)
    v = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert_close(dd, torch.logdet(J), 1e-12)
    assert_close(mm, v @ torch.linalg.solve(J, v), 1e-12)
    # shape errors mirror the reference's assertion (cyclic_reduction.py:223)
    with pytest.raises(AssertionError):
        c.decompose(torch.eye(2).repeat(3, 1, 1), torch.eye(2).repeat(3, 1, 1))
    # non positive definite input raises instead of returning NaNs
    Rbad = torch.eye(2, dtype=torch.float64).repeat(4, 1, 1)
    Rbad[2] = -Rbad[2]
    with pytest.raises(c.NotPositiveDefiniteError):
        c.decompose(Rbad, torch.zeros(3, 2, 2, dtype=torch.float64))
    # keyword call used by the reference model (models.py:290,367)
    R, O, x = leg_inputs(2, 9, torch.float64)
    c.decompose(**{"Rs": R, "Os": O})
    c.mahal_and_det(Rs=R, Os=O, x=x)


def test_plain_tuple_decomp_is_accepted():
    c = cr()
    R, O, x = leg_inputs(3, 40, torch.float64, seed=3)
    ms, Ds, Fs, Gs = orc.factor(R, O)     # a decomposition that did not come from our decompose()
    plain = (ms, Ds, Fs, Gs)
    dec_o = (ms, Ds, Fs, Gs)
    assert_close(c.solve(plain, x), orc.solve(dec_o, x), 1e-10)
    assert_close(c.det(plain), orc.logdet(dec_o), 1e-12)
    Sd, So = c.inverse_blocks(plain)
    sd, so = orc.selected_inverse(dec_o)
    assert_close(Sd, sd, 1e-10)
    assert_close(So, so, 1e-10)


# ------------------------------------------------------------------ full-size, size-independent properties
def _tridiag_matvec(R, O, w):
    out = torch.einsum("bnij,bnj->bni", R, w)
    out[:, 1:] += torch.einsum("bnij,bnj->bni", O, w[:, :-1])
    out[:, :-1] += torch.einsum("bnji,bnj->bni", O, w[:, 1:])
    return out


def test_config2_shape_properties():
    """B x n = 256 x 10^4, l = 8, fp32 (a quarter of BASELINE config 2's batch): residual of the
    solve, linearity of the solve, and gradient identity gx = 2 gm J^{-1} x."""
    c = cr()
    B, n, l = 256, 10000, 8
    gen = torch.Generator(device="cuda").manual_seed(0)
    G, Bm, LLT = orc.leg_params(l, seed=0)
    gaps = (-torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64, device="cuda")) + 0.01)
    from cyclic_gps.synth import leg_precision_blocks
    R, O = leg_precision_blocks(gaps, G.cuda(), Bm.cuda(), LLT.cuda(), torch.float32)
    x = torch.randn((B, n, l), generator=gen, dtype=torch.float32, device="cuda")
    dec = c.decompose(R, O)
    w = c.solve(dec, x)
    res = _tridiag_matvec(R.double(), O.double(), w.double()) - x.double()
    assert float(res.abs().max() / x.abs().max()) < 1e-4
    w2 = c.solve(dec, 2.0 * x)
    assert relerr(w2, 2.0 * w) < 1e-6
    xr = x.clone().requires_grad_(True)
    mm, dd = c.mahal_and_det(R, O, xr)
    mm.sum().backward()
    assert relerr(xr.grad, 2.0 * w) < 1e-4
    assert relerr(mm, (x.double() * w.double()).sum(dim=(1, 2))) < 1e-4
    # three series against the oracle
    for b in (0, 100, 255):
        dec_o = orc.factor(R[b].double().cpu(), O[b].double().cpu())
        assert_close(dd[b], orc.logdet(dec_o), 1e-4)
        assert_close(w[b], orc.solve(dec_o, x[b].double().cpu()), 1e-4)


def test_not_positive_definite_under_autograd():
    """Default = the reference's contract (cyclic_reduction.py:429): raised inside mahal_and_det, with or without
    autograd.  EAGER_PD_CHECK = False (opt-in, bench.py) defers the report to backward()."""
    c = cr()
    R, O, x = (t.cuda() for t in leg_inputs(3, 200, torch.float64, seed=9))
    R[77] = -R[77]
    with pytest.raises(c.NotPositiveDefiniteError):
        c.mahal_and_det(R, O, x)
    with pytest.raises(c.NotPositiveDefiniteError):
        c.mahal_and_det(R.clone().requires_grad_(True), O, x)
    c.EAGER_PD_CHECK = False
    try:
        Rr = R.clone().requires_grad_(True)
        mm, dd = c.mahal_and_det(Rr, O, x)
        with pytest.raises(c.NotPositiveDefiniteError):
            (mm + dd).backward()
    finally:
        c.EAGER_PD_CHECK = True


def test_backward_twice_and_release_flag():
    """backward(retain_graph=True) followed by another backward works like the reference's tape; the opt-in
    RELEASE_FACTORS_AFTER_BACKWARD frees the factors after the first pass and a second one raises."""
    c = cr()
    R, O, x = (t.cuda() for t in leg_inputs(3, 120, torch.float64, seed=4))
    Rr = R.clone().requires_grad_(True)
    mm, dd = c.mahal_and_det(Rr, O, x)
    (mm + dd).backward(retain_graph=True)
    g1 = Rr.grad.clone()
    Rr.grad = None
    (mm + dd).backward()
    assert relerr(Rr.grad, g1) == 0.0
    (ga,) = torch.autograd.grad(c.mahal_and_det(Rr, O, x)[0], Rr)
    assert ga.shape == Rr.shape
    c.RELEASE_FACTORS_AFTER_BACKWARD = True
    try:
        mm, dd = c.mahal_and_det(Rr, O, x)
        (mm + dd).backward(retain_graph=True)
        with pytest.raises(RuntimeError):
            (mm + dd).backward()
    finally:
        c.RELEASE_FACTORS_AFTER_BACKWARD = False


@pytest.mark.parametrize("l,n,dtype", [(3, 257, torch.float64), (8, 500, torch.float64), (4, 333, torch.float32)])
def test_solve_is_differentiable(l, n, dtype):
    """Gradient of a scalar function of solve(decompose(Rs, Os), y) wrt y, Rs, Os against torch autograd through the
    oracle (the reference's own autograd path for the same expression)."""
    c = cr()
    R, O, x = leg_inputs(l, n, dtype, seed=5 * l + n)
    coef = torch.randn((n, l), generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    Rg, Og, xg = (t.cuda().requires_grad_(True) for t in (R, O, x))
    w = c.solve(c.decompose(Rg, Og), xg)
    (w * coef.to(dtype).cuda()).sum().backward()
    Ro, Oo, xo = (t.double().requires_grad_(True) for t in (R, O, x))
    wo = orc.solve(orc.factor(Ro, Oo), xo)
    (wo * coef).sum().backward()
    tol = TOL[dtype]
    assert_close(w, wo, tol, "solve")
    assert_close(xg.grad, xo.grad, tol, "d/dy")
    assert_close(Rg.grad, Ro.grad, tol, "d/dRs")
    assert_close(Og.grad, Oo.grad, tol, "d/dOs")


@pytest.mark.parametrize("l,n,dtype,batched", [(1, 1, torch.float64, False), (2, 2, torch.float64, False), (3, 33, torch.float64, True),
                                                (5, 100, torch.float64, False), (8, 257, torch.float32, True), (16, 40, torch.float64, False)])
def test_inverse_blocks_is_differentiable(l, n, dtype, batched):
    """gradient of a scalar function of inverse_blocks(decompose(Rs, Os)) wrt Rs, Os against torch autograd
    through the oracle (= the reference's autograd path for the same expression, cyclic_reduction.py:470-503).  The forward
    values come from the CUDA kernels, the backward from the torch-op adjoint recursion (cyclic_gps/_adjoint.py)."""
    c = cr()
    tol = TOL[dtype]
    series = [leg_inputs(l, n, dtype, seed=9 * l + n + b) for b in range(3 if batched else 1)]
    gen = torch.Generator().manual_seed(3)
    gR_want, gO_want, Sd_want, So_want, cds, cos = [], [], [], [], [], []
    for (R, O, _) in series:
        Ro, Oo = R.double().clone().requires_grad_(True), O.double().clone().requires_grad_(True)
        Sd0, So0 = orc.selected_inverse(orc.factor(Ro, Oo))
        cd, co = torch.randn(Sd0.shape, generator=gen, dtype=torch.float64), torch.randn(So0.shape, generator=gen, dtype=torch.float64)
        ((Sd0 * cd).sum() + (So0 * co).sum()).backward()
        gR_want.append(Ro.grad); gO_want.append(Oo.grad if n > 1 else torch.zeros_like(Oo)); Sd_want.append(Sd0.detach()); So_want.append(So0.detach())
        cds.append(cd); cos.append(co)
    st = (lambda xs: torch.stack(xs)) if batched else (lambda xs: xs[0])
    Rg = st([s[0] for s in series]).cuda().requires_grad_(True)
    Og = st([s[1] for s in series]).cuda().requires_grad_(True)
    Sd, So = c.inverse_blocks(c.decompose(Rg, Og))
    assert Sd.requires_grad
    ((Sd * st(cds).to(dtype).cuda()).sum() + (So * st(cos).to(dtype).cuda()).sum()).backward()
    assert_close(Sd, st(Sd_want), tol, "Sigma_d")
    assert_close(Rg.grad, st(gR_want), 10 * tol, "d/dRs")
    if n > 1:
        assert_close(So, st(So_want), tol, "Sigma_o")
        assert_close(Og.grad, st(gO_want), 10 * tol, "d/dOs")
    # no grad requested: the plain kernel path, nothing attached
    Sd2, So2 = c.inverse_blocks(c.decompose(Rg.detach(), Og.detach()))
    assert not Sd2.requires_grad
    assert_close(Sd2, Sd.detach(), 0.0 if dtype == torch.float64 else tol, "same forward")


@pytest.mark.parametrize("l,n,dtype,batched", [(1, 1, torch.float64, False), (3, 33, torch.float64, True), (8, 1000, torch.float32, True),
                                                (16, 257, torch.float64, False), (5, 2, torch.float64, False)])
def test_solve_and_inverse_blocks_is_the_reference_sequence(l, n, dtype, batched):
    """The in-sample posterior in two sweeps (forward with the right-hand side, backward with back-solve + selected inverse)
    against decompose -> solve -> inverse_blocks (reference models.py:282-298) and against the oracle."""
    c = cr()
    tol = TOL[dtype]
    series = [leg_inputs(l, n, dtype, seed=3 * l + n + b) for b in range(3 if batched else 1)]
    st = (lambda xs: torch.stack(xs)) if batched else (lambda xs: xs[0])
    R, O, x = (st([s[i] for s in series]).cuda() for i in range(3))
    w, Sd, So = c.solve_and_inverse_blocks(R, O, x)
    dec = c.decompose(R, O)
    w2 = c.solve(dec, x)
    Sd2, So2 = c.inverse_blocks(dec)
    assert_close(w, w2, tol, "mean vs solve")
    assert_close(Sd, Sd2, tol, "Sigma_d vs inverse_blocks")
    if n > 1:
        assert_close(So, So2, tol, "Sigma_o vs inverse_blocks")
    for b, (Rb, Ob, xb) in enumerate(series):
        dec_o = orc.factor(Rb.double(), Ob.double())
        sd, so = orc.selected_inverse(dec_o)
        pick = (lambda t: t[b]) if batched else (lambda t: t)
        assert_close(pick(w), orc.solve(dec_o, xb.double()), tol, "mean vs oracle")
        assert_close(pick(Sd), sd, tol, "Sigma_d vs oracle")
        if n > 1:
            assert_close(pick(So), so, tol, "Sigma_o vs oracle")


def test_check_decompose_loop_outputs_like_the_reference():
    """reference cyclic_reduction.py:262-280: the shape check of one level (AssertionError on a mismatch)."""
    c = cr()
    for n in (2, 5, 8):
        R, O, _ = leg_inputs(3, n, torch.float64, seed=n)
        (m, K, F, G), (Rn, On) = c.decompose_step(R, O)
        c.check_decompose_loop_outputs(m, K, F, G, Rn, On)
        with pytest.raises(AssertionError):
            c.check_decompose_loop_outputs(m + 2, K, F, G, Rn, On)


@pytest.mark.parametrize("l,n,dtype,batch", [(3, 1000, torch.float64, None), (8, 700, torch.float32, 6), (16, 300, torch.float64, 2)])
def test_graphed_mahal_and_det_replays(l, n, dtype, batch):
    """cyclic_gps.graphs.GraphedMahalAndDet: a CUDA-graph replay gives the numbers of the eager path, for new inputs too."""
    from cyclic_gps.graphs import GraphedMahalAndDet
    c = cr()
    tol = TOL[dtype]
    def inputs(seed):
        if batch is None:
            return [t.cuda() for t in leg_inputs(l, n, dtype, seed=seed)]
        parts = [leg_inputs(l, n, dtype, seed=seed + b) for b in range(batch)]
        return [torch.stack([p[i] for p in parts]).cuda() for i in range(3)]
    R, O, x = inputs(1)
    g = GraphedMahalAndDet(R, O, x, g_mahal=0.7, g_det=-1.3)
    for seed in (1, 50, 99):
        R, O, x = inputs(seed)
        mh, ld, gR, gO, gx = g(R, O, x)
        Rr, Or, xr = R.clone().requires_grad_(True), O.clone().requires_grad_(True), x.clone().requires_grad_(True)
        mm, dd = c.mahal_and_det(Rr, Or, xr)
        (0.7 * mm.sum() - 1.3 * dd.sum()).backward()
        for a, b_, name in ((mh, mm, "mahal"), (ld, dd, "logdet"), (gR, Rr.grad, "gR"), (gO, Or.grad, "gO"), (gx, xr.grad, "gx")):
            assert_close(a, b_, tol, f"graphed {name} seed {seed}")
        g.check()


def test_jitter_ladder_like_psd_safe_cholesky():
    """Error-recovery parity with the reference's psd_safe_cholesky calls (cyclic_reduction.py:227,306,429; gpytorch's ladder
    jitter * 10**i on the whole batch of a level): a barely indefinite even block is repaired by the first rung, a badly
    indefinite one raises NotPositiveDefiniteError (= NotPSDError) after three warnings, NaNs raise NanError."""
    import warnings
    c = cr()
    R, O, x = leg_inputs(3, 40, torch.float64, seed=11)
    eye = torch.eye(3, dtype=torch.float64)
    # (i) a single block whose smallest eigenvalue is -3e-9: the factorisation fails, the first rung (+1e-8) repairs it
    R1 = R[4:5] - (torch.linalg.eigvalsh(R[4]).min() + 3e-9) * eye
    O1 = O[:0]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mm, dd = c.mahal_and_det(R1.cuda(), O1.cuda(), x[:1].cuda())
        dec = c.decompose(R1.cuda(), O1.cuda())
    assert sum(issubclass(i.category, c.NumericalWarning) for i in w) == 2            # one rung per call
    d_o = orc.factor(R1 + 1e-8 * eye, O1)                                              # what psd_safe_cholesky factorises
    assert_close(mm, orc.mahal(d_o, x[:1]), 1e-6, "mahal after jitter")
    assert_close(dd, orc.logdet(d_o), 1e-7, "logdet after jitter")
    assert_close(dec[1][0], d_o[1][0], 1e-6, "D[0] after jitter")
    # (ii) one level of a longer system: the jitter goes onto EVERY even block of the level, as in the reference's batched call
    R3 = R[:3].clone()
    R3[2] -= (torch.linalg.eigvalsh(R3[2]).min() + 3e-9) * eye
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        (m3, K, F, G), (Rn, On) = c.decompose_step(R3.cuda(), O[:2].cuda())
    assert sum(issubclass(i.category, c.NumericalWarning) for i in w) == 1
    Rj = R3.clone()
    Rj[0::2] += 1e-8 * eye
    (_, K2, F2, G2), (Rn2, _) = orc.level_step(Rj, O[:2])
    assert_close(K[0], K2[0], 1e-12, "K of the healthy even block carries the jitter too")
    assert_close(F, F2, 1e-12, "F")
    assert_close(G, G2, 1e-5, "G (through the repaired block)")
    # hopeless: three rungs, then the error (at level 1: the bad block is an odd row of level 0)
    Rc = R.clone()
    Rc[1] -= 10.0 * torch.eye(3, dtype=torch.float64)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(c.NotPSDError, match="level 1.*jitter up to 1.0e-06"):
            c.mahal_and_det(Rc.cuda(), O.cuda(), x.cuda())
    assert sum(issubclass(i.category, c.NumericalWarning) for i in w) == 3
    # fp32 default ladder starts at 1e-6; JITTER overrides it; CHOLESKY_MAX_TRIES = 0 switches the retry off
    c.CHOLESKY_MAX_TRIES = 0
    try:
        with pytest.raises(c.NotPositiveDefiniteError):
            c.mahal_and_det(R1.cuda(), O1.cuda(), x[:1].cuda())
    finally:
        c.CHOLESKY_MAX_TRIES = 3
    Rn = R.clone()
    Rn[6, 0, 0] = float("nan")
    with pytest.raises(c.NanError):
        c.decompose(Rn.cuda(), O.cuda())
