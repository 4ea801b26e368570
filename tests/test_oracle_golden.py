"""CPU: pins oracle/cr_oracle.py to the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), including the reference's own known-answer
cases (tests/test_cyclic_reduction.py:147-291 of the reference)."""
import numpy as np
import pytest
import torch

from helpers import assert_close, split_levels
from oracle import cr_oracle as orc


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _check_case(g, p, tol, factors=True):
    R, O, x = _t(g[p + "R"]), _t(g[p + "O"]), _t(g[p + "x"])
    l = R.shape[1]
    dec = orc.factor(R, O)
    assert np.array_equal(dec[0].numpy(), g[p + "ms"])
    if factors and (p + "D") in g.files:
        for name, lv in (("D", dec[1]), ("F", dec[2]), ("G", dec[3])):
            want = split_levels(g[p + name], g[p + name + "_counts"], (l, l))
            assert len(want) == len(lv)
            for a, b in zip(lv, want):
                assert_close(a, b, tol, p + name)
    hs = orc.forward_sub(dec, x)
    want = split_levels(g[p + "halfsolve"], g[p + "halfsolve_counts"], (l,))
    for a, b in zip(hs, want):
        assert_close(a, b, tol, p + "halfsolve")
    assert_close(orc.solve(dec, x), g[p + "solve"], tol, p + "solve")
    assert_close(orc.logdet(dec), g[p + "logdet"], tol, p + "logdet")
    assert_close(orc.mahal(dec, x), g[p + "mahal"], tol, p + "mahal")
    mm, dd = orc.mahal_and_logdet(R, O, x)
    assert_close(mm, g[p + "mahal_fused"], tol, p + "mahal_fused")
    assert_close(dd, g[p + "logdet_fused"], tol, p + "logdet_fused")
    Sd, So = orc.selected_inverse(dec)
    assert_close(Sd, g[p + "Sd"], tol, p + "Sd")
    assert_close(So, g[p + "So"], tol, p + "So")
    counts = [(int(m) + 1) // 2 for m in g[p + "ms"]]
    ycrr = split_levels(g[p + "ycrr"], counts, (l,))
    assert_close(orc.backward_sub(dec, ycrr), g[p + "backhalfsolve"], tol, p + "backhalfsolve")
    for tag, (gm, gd) in (("11", (1.0, 1.0)), ("ab", (0.3, -0.7))):
        gR, gO, gx = orc.loglik_grads(R, O, x, gm, gd)
        gtol = tol * 10
        assert_close(gR, g[p + "gR_" + tag], gtol, p + "gR")
        assert_close(gO, g[p + "gO_" + tag], gtol, p + "gO")
        assert_close(gx, g[p + "gx_" + tag], gtol, p + "gx")


def test_oracle_random_llt(golden):
    g = golden["random_llt"]
    for p in g["cases"]:
        p = str(p)
        _check_case(g, p, 1e-9)   # the reference's J=LL^T generator is ill-conditioned for n~33
        R, O, x = _t(g[p + "R"]), _t(g[p + "O"]), _t(g[p + "x"])
        dec = orc.factor(R, O)
        assert_close(orc.logdet(dec), g[p + "dense_logdet"], 1e-9, p + "dense logdet")
        assert_close(orc.solve(dec, x), g[p + "dense_solve"], 1e-6, p + "dense solve")


def test_oracle_leg(golden):
    g = golden["leg"]
    for p in g["cases"]:
        p = str(p)
        tol = 2e-5 if "float32" in p else 1e-11
        _check_case(g, p, tol)


def test_oracle_known_matrices(golden):
    g = golden["known"]
    x = _t(g["x"])
    for name, l in (("bab", 1), ("schur", 2)):
        R, O = _t(g[name + "_R"]), _t(g[name + "_O"])
        dec = orc.factor(R, O)
        xv = x.reshape(-1, l)
        mm, dd = orc.mahal_and_logdet(R, O, xv)
        # closed forms, with the reference test's np.allclose defaults (rtol 1e-5, atol 1e-8)
        assert np.allclose(g[name + "_logdet_closed"], orc.logdet(dec).numpy())
        assert np.allclose(g[name + "_logdet_closed"], dd.numpy())
        assert np.allclose(g[name + "_mahal_closed"], mm.numpy())
        Sd, So = orc.selected_inverse(dec)
        assert np.allclose(g[name + "_inv_R_closed"], Sd.numpy())
        assert np.allclose(g[name + "_inv_O_closed"], So.numpy())
        # and the reference's own CR outputs
        assert_close(Sd, g[name + "_inv_R_ref"], 1e-5, name)
        assert_close(So, g[name + "_inv_O_ref"], 1e-5, name)
        assert_close(mm, g[name + "_mahal_ref"], 1e-5, name)
        assert_close(dd, g[name + "_logdet_fused_ref"], 1e-5, name)


def test_oracle_helpers(golden):
    g = golden["helpers"]
    for p in g["cases"]:
        p = str(p)
        F, G, x, y, Sd, So = (_t(g[p + k]) for k in ("F", "G", "x", "y", "Sd", "So"))
        a, b = orc.bidiag_gram(F, G)
        assert_close(a, g[p + "UUT_d"], 1e-13, p)
        assert_close(b, g[p + "UUT_o"], 1e-13, p)
        assert_close(orc.bidiag_mv(F, G, x), g[p + "Ux"], 1e-13, p)
        assert_close(orc.bidiag_tmv(F, G, y), g[p + "UTx"], 1e-13, p)
        a, b = orc.symtri_times_bidiag(Sd, So, F, G)
        assert_close(a, g[p + "SigU_d"], 1e-13, p)
        assert_close(b, g[p + "SigU_o"], 1e-13, p)
        assert_close(orc.bidiag_t_bidiag_diag(F, G, a, b), g[p + "UtV"], 1e-13, p)
    for key in g.files:
        if key.startswith("il_") and key.endswith("_out"):
            base = key[:-4]
            out = orc.interleave(_t(g[base + "_a"]), _t(g[base + "_b"]))
            assert np.array_equal(out.numpy(), g[key])


def test_oracle_autograd_equals_closed_form():
    torch.manual_seed(0)
    G, B, LLT = orc.leg_params(4, seed=3)
    gaps = torch.rand(40, dtype=torch.float64) + 0.05
    R, O = orc.leg_posterior_precision(gaps, G, B, LLT)
    x = torch.randn(41, 4, dtype=torch.float64)
    (_, _), (aR, aO, ax) = orc.loglik_grads_autograd(R, O, x, 0.5, 2.0)
    cR, cO, cx = orc.loglik_grads(R, O, x, 0.5, 2.0)
    assert_close(cR, aR, 1e-12)
    assert_close(cO, aO, 1e-12)
    assert_close(cx, ax, 1e-12)


def test_oracle_elimination_order():
    assert orc.elimination_order(6) == [0, 2, 4, 1, 5, 3]
    assert orc.elimination_order(1) == [0]
    # the dense Cholesky of the permuted matrix reproduces forward_sub (reference test :171-191)
    torch.manual_seed(1)
    n, l = 13, 2
    G, B, LLT = orc.leg_params(l, seed=1)
    R, O = orc.leg_posterior_precision(torch.rand(n - 1, dtype=torch.float64) + 0.1, G, B, LLT)
    J = orc.assemble_dense(R, O)
    perm = orc.elimination_order(n)
    idx = torch.tensor([[p * l + a for a in range(l)] for p in perm]).reshape(-1)
    Lp = torch.linalg.cholesky(J[idx][:, idx])
    v = torch.randn(n, l, dtype=torch.float64)
    want = torch.linalg.solve_triangular(Lp, v.reshape(-1)[idx].unsqueeze(-1), upper=False).squeeze(-1)
    got = torch.cat(orc.forward_sub(orc.factor(R, O), v)).reshape(-1)
    assert_close(got, want, 1e-12)
