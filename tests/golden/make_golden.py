"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is imported through the stand-in modules in oracle/ref_shims (gpytorch,
torchtyping, pytorch_lightning, filterpy are not installed; typeguard 4 breaks the
reference's @typechecked functions, so the decorator is made the identity).  Everything
stored is numeric input/output data of the reference's own functions:

  random_llt.npz     random block-bidiagonal J = L L^T cases of
                     tests/test_cyclic_reduction.py:147-223 (seeded), every CR function
  known.npz          BAB / Schur-block closed forms (tests/test_cyclic_reduction.py:243-291,
                     tests/known_matrices_full.py) in fp32
  leg.npz            LEG posterior-precision inputs: scalars, solves, selected inverse,
                     factors and reference-autograd gradients
  helpers.npz        UU_T / Ux / U_Tx / SigU / UtV_diags / interleave
  leg_model.npz      LEGFamily.log_likelihood / compute_insample_posterior values
  predictions.npz    LEGFamily.predictive_posterior / make_predictions values (models.py:394-546)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, "/root/reference/tests")
sys.path.insert(3, ROOT)

import typeguard  # noqa: E402

typeguard.typechecked = lambda f=None, **k: f if f is not None else (lambda g: g)

from cyclic_gps import cyclic_reduction as ref  # noqa: E402  (the reference)
from cyclic_gps.models import LEGFamily  # noqa: E402
from cyclic_gps.model_utils import compute_log_marginal_likelihood  # noqa: E402
from oracle import cr_oracle as orc  # noqa: E402  (only for the LEG input generator)

assert ref.__file__.startswith("/root/reference"), ref.__file__


def npy(t):
    return t.detach().cpu().numpy()


def cat_levels(levels):
    """list of per-level tensors -> one flat array + row counts."""
    counts = np.array([int(t.shape[0]) for t in levels], dtype=np.int64)
    if len(levels) == 0:
        return np.zeros((0,)), counts
    width = int(np.prod(levels[0].shape[1:]))
    flat = np.concatenate([npy(t).reshape(t.shape[0], width) for t in levels], axis=0)
    return flat, counts


def all_outputs(R, O, x, prefix, out, with_factors=True, with_grads=True, ycrr_seed=None):
    dec = ref.decompose(R, O)
    ms, Ds, Fs, Gs = dec
    out[prefix + "R"], out[prefix + "O"], out[prefix + "x"] = npy(R), npy(O), npy(x)
    out[prefix + "ms"] = npy(ms)
    if with_factors:
        for name, lv in (("D", Ds), ("F", Fs), ("G", Gs)):
            out[prefix + name], out[prefix + name + "_counts"] = cat_levels(lv)
    hs = ref.halfsolve(dec, x)
    out[prefix + "halfsolve"], out[prefix + "halfsolve_counts"] = cat_levels(hs)
    out[prefix + "solve"] = npy(ref.solve(dec, x))
    out[prefix + "logdet"] = npy(ref.det(dec)) if R.shape[0] <= 2000 else npy(ref.mahal_and_det(R, O, x)[1])
    out[prefix + "mahal"] = npy(ref.mahal(dec, x))
    m2, d2 = ref.mahal_and_det(R, O, x)
    out[prefix + "mahal_fused"], out[prefix + "logdet_fused"] = npy(m2), npy(d2)
    Sd, So = ref.inverse_blocks(dec)
    out[prefix + "Sd"], out[prefix + "So"] = npy(Sd), npy(So)
    if ycrr_seed is not None:
        g = torch.Generator().manual_seed(ycrr_seed)
        ycrr = [torch.randn(((int(m) + 1) // 2, R.shape[1]), generator=g, dtype=torch.float64).to(R.dtype) for m in ms]
        out[prefix + "ycrr"], _ = cat_levels(ycrr)
        out[prefix + "backhalfsolve"] = npy(ref.backhalfsolve(dec, ycrr))
    if with_grads:
        for tag, (gm, gd) in (("11", (1.0, 1.0)), ("ab", (0.3, -0.7))):
            Rr, Or, xr = [t.clone().requires_grad_(True) for t in (R, O, x)]
            mm, dd = ref.mahal_and_det(Rr, Or, xr)
            (gm * mm + gd * dd).backward()
            out[prefix + "gR_" + tag], out[prefix + "gO_" + tag], out[prefix + "gx_" + tag] = (
                npy(Rr.grad), npy(Or.grad if Or.grad is not None else torch.zeros_like(Or)), npy(xr.grad))


def gen_random_llt():
    rng = np.random.RandomState(10)
    out = {}
    cases = []
    for l in (1, 3):
        for n in (2, 6, 30, 31, 32, 33):
            Ld = rng.randn(n, l, l) + 3 * np.eye(l)
            Lo = rng.randn(n - 1, l, l)
            L = np.zeros((n, l, n, l))
            for i in range(n):
                L[i, :, i] = Ld[i]
            for i in range(1, n):
                L[i, :, i - 1] = Lo[i - 1]
            L = L.reshape(n * l, n * l)
            J = (L @ L.T).reshape(n, l, n, l)
            R = torch.from_numpy(np.array([J[i, :, i] for i in range(n)]))
            O = torch.from_numpy(np.array([J[i, :, i - 1] for i in range(1, n)]))
            x = torch.from_numpy(rng.randn(n, l))
            p = f"l{l}_n{n}_"
            all_outputs(R, O, x, p, out, ycrr_seed=100 + n)
            Jd = J.reshape(n * l, n * l)
            out[p + "dense_logdet"] = np.linalg.slogdet(Jd)[1]
            out[p + "dense_solve"] = np.linalg.solve(Jd, npy(x).ravel()).reshape(n, l)
            cases.append(p)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "random_llt.npz"), **out)


def block_rep(M, l):
    n = M.shape[0] // l
    R = torch.empty((n, l, l))
    O = torch.empty((n - 1, l, l))
    for i in range(n):
        R[i] = M[i * l:(i + 1) * l, i * l:(i + 1) * l]
    for i in range(n - 1):
        O[i] = M[(i + 1) * l:(i + 2) * l, i * l:(i + 1) * l]
    return R, O


def gen_known():
    from known_matrices_full import (bab_determinant, bab_inverse, bab_matrix, schur_block_determinant,
                                     schur_block_inverse, schur_block_matrix)
    out = {}
    g = torch.Generator().manual_seed(3)
    x = torch.rand((10, 1), generator=g)
    out["x"] = npy(x)
    # BAB(10, 5, 2), block size 1   (test_cyclic_reduction.py:246-266)
    BAB = torch.from_numpy(bab_matrix(n=10, alpha=5, beta=2))
    R, O = block_rep(BAB, 1)
    out["bab_R"], out["bab_O"] = npy(R), npy(O)
    out["bab_logdet_closed"] = np.log(bab_determinant(10, 5, 2))
    inv = torch.from_numpy(bab_inverse(10, 5, 2)).float()
    iR, iO = block_rep(inv, 1)
    out["bab_inv_R_closed"], out["bab_inv_O_closed"] = npy(iR), npy(iO)
    out["bab_mahal_closed"] = npy(x.T @ inv @ x)
    dec = ref.decompose(R, O)
    out["bab_logdet_ref"] = npy(ref.det(dec))
    mm, dd = ref.mahal_and_det(R, O, x=x)
    out["bab_mahal_ref"], out["bab_logdet_fused_ref"] = npy(mm), npy(dd)
    sR, sO = ref.inverse_blocks(dec)
    out["bab_inv_R_ref"], out["bab_inv_O_ref"] = npy(sR), npy(sO)
    # Gram of Schur-block(10, 1.., 2..), block size 2   (:268-291)
    S = torch.from_numpy(schur_block_matrix(n=10, x=[1] * 10, y=[2] * 9))
    S = S.T @ S
    R, O = block_rep(S, 2)
    out["schur_R"], out["schur_O"] = npy(R), npy(O)
    out["schur_logdet_closed"] = np.log(schur_block_determinant(n=10, x=[1] * 10, y=[2] * 9) ** 2)
    inv = torch.from_numpy(schur_block_inverse(n=10, x=[1] * 10, y=[2] * 9)).float()
    inv = inv @ inv.T
    iR, iO = block_rep(inv, 2)
    out["schur_inv_R_closed"], out["schur_inv_O_closed"] = npy(iR), npy(iO)
    out["schur_mahal_closed"] = npy(x.T @ inv @ x)
    dec = ref.decompose(R, O)
    out["schur_logdet_ref"] = npy(ref.det(dec))
    mm, dd = ref.mahal_and_det(R, O, x=x.reshape(5, 2))
    out["schur_mahal_ref"], out["schur_logdet_fused_ref"] = npy(mm), npy(dd)
    sR, sO = ref.inverse_blocks(dec)
    out["schur_inv_R_ref"], out["schur_inv_O_ref"] = npy(sR), npy(sO)
    np.savez_compressed(os.path.join(HERE, "known.npz"), **out)


LEG_CASES = [  # (rank, n, dtype, spacing, keep factors)
    (3, 1000, "float64", "irregular", False),   # BASELINE config 1
    (2, 64, "float64", "regular", True),
    (4, 257, "float32", "irregular", True),
    (5, 33, "float64", "irregular", True),
    (8, 100, "float32", "irregular", False),
    (8, 65, "float64", "irregular", True),
    (16, 37, "float64", "gap", False),          # CO2-shaped: one long gap mid-series
    (1, 17, "float64", "irregular", True),
    (3, 1, "float64", "regular", True),
]


def leg_case_inputs(rank, n, dtype, spacing, seed):
    g = torch.Generator().manual_seed(seed)
    if spacing == "regular":
        gaps = torch.ones(max(n - 1, 0), dtype=torch.float64)
    else:
        gaps = -torch.log(torch.rand(max(n - 1, 0), generator=g, dtype=torch.float64)) + 0.01
        if spacing == "gap":
            gaps[n // 2] = 240.0 / 12.0
    G, B, LLT = orc.leg_params(rank, seed=seed)
    if n == 1:
        R = (torch.eye(rank, dtype=torch.float64) + B.T @ torch.linalg.solve(LLT, B)).unsqueeze(0)
        O = torch.zeros((0, rank, rank), dtype=torch.float64)
    else:
        R, O = orc.leg_posterior_precision(gaps, G, B, LLT)
    x = torch.randn((n, rank), generator=g, dtype=torch.float64)
    dt = getattr(torch, dtype)
    return R.to(dt), O.to(dt), x.to(dt)


def gen_leg():
    out = {}
    cases = []
    for i, (rank, n, dtype, spacing, keep) in enumerate(LEG_CASES):
        R, O, x = leg_case_inputs(rank, n, dtype, spacing, seed=1000 + i)
        p = f"l{rank}_n{n}_{dtype}_"
        all_outputs(R, O, x, p, out, with_factors=keep, ycrr_seed=7 + i)
        cases.append(p)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "leg.npz"), **out)


def gen_helpers():
    rng = np.random.RandomState(5)
    out = {}
    cases = []
    for (l, o, wide) in [(1, 4, True), (1, 4, False), (2, 3, True), (2, 3, False), (3, 1, True), (5, 6, False), (5, 6, True)]:
        F = torch.from_numpy(rng.randn(o, l, l))
        G = torch.from_numpy(rng.randn(o if wide else o - 1, l, l))
        x = torch.from_numpy(rng.randn(o + 1 if wide else o, l))
        y = torch.from_numpy(rng.randn(o, l))
        S = rng.randn(o * l, o * l)
        S = (S @ S.T).reshape(o, l, o, l)
        Sd = torch.from_numpy(np.array([S[i, :, i] for i in range(o)]))
        So = torch.from_numpy(np.array([S[i + 1, :, i] for i in range(o - 1)]).reshape(o - 1, l, l))
        p = f"l{l}_o{o}_{'wide' if wide else 'sq'}_"
        out[p + "F"], out[p + "G"], out[p + "x"], out[p + "y"], out[p + "Sd"], out[p + "So"] = map(npy, (F, G, x, y, Sd, So))
        a, b = ref.UU_T(F, G)
        out[p + "UUT_d"], out[p + "UUT_o"] = npy(a), npy(b)
        out[p + "Ux"] = npy(ref.Ux(F, G, x))
        out[p + "UTx"] = npy(ref.U_Tx(F, G, y))
        a, b = ref.SigU(Sd, So, F, G)
        out[p + "SigU_d"], out[p + "SigU_o"] = npy(a), npy(b)
        out[p + "UtV"] = npy(ref.UtV_diags(F, G, a, b))
        cases.append(p)
    for (na, nb) in [(3, 3), (4, 3), (3, 4), (3, 5), (1, 0)]:
        a = torch.from_numpy(rng.randn(na, 2))
        b = torch.from_numpy(rng.randn(nb, 2))
        out[f"il_{na}_{nb}_a"], out[f"il_{na}_{nb}_b"] = npy(a), npy(b)
        out[f"il_{na}_{nb}_out"] = npy(ref.interleave(a, b))
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), **out)


def gen_leg_model():
    out = {}
    cases = []
    torch.manual_seed(11)
    for spacing in ("regular", "irregular"):
        for n in (10, 33):
            for d in (1, 2):
                g = torch.Generator().manual_seed(n * 10 + d)
                if spacing == "regular":
                    ts = torch.arange(n, dtype=torch.float64)
                else:
                    ts = torch.cumsum(-torch.log(torch.rand(n, generator=g, dtype=torch.float64)) + 0.01, 0)
                xs = torch.randn((n, d), generator=g, dtype=torch.float64)
                model = LEGFamily(rank=5, obs_dim=d, train=True, data_type=torch.float64)
                model.double()
                p = f"{spacing}_n{n}_d{d}_"
                out[p + "ts"], out[p + "xs"] = npy(ts), npy(xs)
                for name in ("N_params", "R_params", "Lambda_params", "B"):
                    out[p + name] = npy(getattr(model, name))
                ll = model.log_likelihood(ts=ts, xs=xs)
                out[p + "ll"] = npy(ll)
                ll.backward()
                for name in ("N_params", "R_params", "Lambda_params", "B"):
                    out[p + "grad_" + name] = npy(getattr(model, name).grad)
                out[p + "ll_dense"] = npy(compute_log_marginal_likelihood(
                    N=model.N, R=model.R, B=model.B, Lambda=model.calc_Lambda_Lambda_T(model.Lambda), ts=ts, xs=xs))
                with torch.no_grad():
                    mean, cov = model.compute_insample_posterior(ts, xs)
                out[p + "post_mean"], out[p + "post_cov_R"], out[p + "post_cov_O"] = npy(mean), npy(cov["Rs"]), npy(cov["Os"])
                cases.append(p)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "leg_model.npz"), **out)


def gen_predictions():
    """LEGFamily.make_predictions / predictive_posterior of the reference (models.py:394-546): targets before the
    first observation, at the first, between observations, AT an inner observation time, at the last and beyond it."""
    out = {}
    cases = []
    torch.manual_seed(23)
    for spacing in ("regular", "irregular"):
        for (n, d, rank) in ((12, 1, 3), (40, 2, 5)):
            g = torch.Generator().manual_seed(7 * n + d)
            if spacing == "regular":
                ts = torch.arange(n, dtype=torch.float64)
            else:
                ts = torch.cumsum(-torch.log(torch.rand(n, generator=g, dtype=torch.float64)) + 0.05, 0)
            xs = torch.randn((n, d), generator=g, dtype=torch.float64)
            inner = ts[:-1] + (ts[1:] - ts[:-1]) * torch.rand(n - 1, generator=g, dtype=torch.float64)
            target = torch.cat([ts[:1] - 2.5, ts[:1] - 0.3, ts[:1], inner[::3], ts[n // 2:n // 2 + 1], ts[-1:], ts[-1:] + 0.4, ts[-1:] + 3.0])
            target = torch.sort(torch.unique(target))[0]
            model = LEGFamily(rank=rank, obs_dim=d, train=False, data_type=torch.float64)
            model.double()
            p = f"{spacing}_n{n}_d{d}_r{rank}_"
            out[p + "ts"], out[p + "xs"], out[p + "target"] = npy(ts), npy(xs), npy(target)
            for name in ("N_params", "R_params", "Lambda_params", "B"):
                out[p + name] = npy(getattr(model, name))
            with torch.no_grad():
                zm, zv = model.predictive_posterior(ts, xs, target)
                xm, xv = model.make_predictions(ts, xs, target)
            out[p + "z_mean"], out[p + "z_cov"], out[p + "x_mean"], out[p + "x_cov"] = npy(zm), npy(zv), npy(xm), npy(xv)
            cases.append(p)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "predictions.npz"), **out)


if __name__ == "__main__":
    if "--predictions-only" in sys.argv:      # added after the other files were frozen
        gen_predictions()
        sys.exit(0)
    gen_random_llt()
    gen_known()
    gen_leg()
    gen_helpers()
    gen_leg_model()
    gen_predictions()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
