"""CPU: the LEG glue around the hot path (dense oracle likelihood, Kalman comparator, precision
builder) against the golden values of the unmodified reference.  No CR call is made here."""
import numpy as np
import torch

from helpers import assert_close
from oracle import cr_oracle as orc


def _model(g, p, d):
    from cyclic_gps.models import LEGFamily
    m = LEGFamily(rank=5, obs_dim=d, train=False, data_type=torch.float64)
    for name in ("N_params", "R_params", "Lambda_params", "B"):
        getattr(m, name).data = torch.from_numpy(g[p + name].copy())
    m.register_model_matrices_from_params()
    return m


def test_dense_likelihood_and_kalman_match_reference(golden):
    from cyclic_gps.kalman import init_kalman_filter, kf_log_marginal_likelihood
    from cyclic_gps.model_utils import compute_log_marginal_likelihood
    g = golden["leg_model"]
    for p in g["cases"]:
        p = str(p)
        d = int(p.split("_d")[1][0])
        m = _model(g, p, d)
        ts, xs = torch.from_numpy(g[p + "ts"]), torch.from_numpy(g[p + "xs"])
        dense = compute_log_marginal_likelihood(N=m.N, R=m.R, B=m.B, Lambda=m.calc_Lambda_Lambda_T(m.Lambda), ts=ts, xs=xs)
        assert_close(dense.reshape(()), g[p + "ll_dense"].reshape(()), 1e-11, p)
        if p.startswith("regular"):
            kf = init_kalman_filter(leg_model=m, use_approximation=False)
            assert_close(torch.tensor(kf_log_marginal_likelihood(kf, xs.numpy())), g[p + "ll"], 1e-6, p + "kalman")   # reference test: torch.allclose defaults


def test_precision_blocks_and_oracle_likelihood(golden):
    """The model's precision builder feeds the oracle CR to the reference's log-likelihood."""
    import math
    g = golden["leg_model"]
    for p in g["cases"]:
        p = str(p)
        d = int(p.split("_d")[1][0])
        m = _model(g, p, d)
        ts, xs = torch.from_numpy(g[p + "ts"]), torch.from_numpy(g[p + "xs"])
        with torch.no_grad():
            Rs, Os = m.compute_PEG_precision(ts)
            KR, KO = m.compute_posterior_precision(ts)
            LLT = m.calc_Lambda_Lambda_T(m.Lambda)
            white = torch.linalg.solve(LLT, xs.T).T
            v = m.compute_v(xs)
            mh, ld = orc.mahal_and_logdet(KR, KO, v)
            prior = orc.logdet(orc.factor(Rs, Os))
            ll = -0.5 * ((torch.sum(white * xs) - mh) + (torch.logdet(2 * math.pi * LLT) * xs.shape[0] + ld - prior))
        assert_close(ll, g[p + "ll"], 1e-10, p)


def test_generate_data_shapes():
    from cyclic_gps.data_utils import generate_data, time_series_dataset
    ts, xs = generate_data(num_datapoints=20, data_dim=3, data_type=torch.float64, spacing="irregular")
    assert ts.shape == (20,) and xs.shape == (20, 3) and bool((ts[1:] > ts[:-1]).all())
    ds = time_series_dataset(ts.unsqueeze(0), xs.unsqueeze(0))
    assert len(ds) == 1 and ds[0][0].shape == (20,)


def test_intercast_vectorised_matches_reference_loop(golden):
    """The vectorised forecast / interpolate / intercast glue (torch only) against golden vectors of the reference's
    per-target Python loop (models.py:454-515); the in-sample posterior comes from the CPU oracle here, the GPU test
    tests/test_likelihood_gpu.py::test_predictions_match_the_reference runs the whole path through the CUDA engine."""
    import torch
    from cyclic_gps.models import LEGFamily
    from oracle import cr_oracle as orc
    g = golden["predictions"]
    for p in [str(c) for c in g["cases"]]:
        ts, xs, target = (torch.from_numpy(g[p + k]) for k in ("ts", "xs", "target"))
        rank = int(p.split("_r")[-1].rstrip("_"))
        m = LEGFamily(rank=rank, obs_dim=xs.shape[-1], train=False, data_type=torch.float64)
        for name in ("N_params", "R_params", "Lambda_params", "B"):
            getattr(m, name).data = torch.from_numpy(g[p + name])
        m.register_model_matrices_from_params()
        with torch.no_grad():
            Rs, Os = m.compute_posterior_precision(ts)
            dec = orc.factor(Rs, Os)
            mean = orc.solve(dec, m.compute_v(xs))
            sd, so = orc.selected_inverse(dec)
            zm, zv = m.intercast(mean, {"Rs": sd, "Os": so}, ts, target)
        for ours, key in ((zm, "z_mean"), (zv, "z_cov")):
            ref = torch.from_numpy(g[p + key])
            assert float((ours - ref).abs().max() / ref.abs().max()) < 1e-10, (p, key)


def test_adjoint_recursion_of_inverse_blocks_matches_oracle_autograd():
    """cyclic_gps/_adjoint.py (torch-op restatement of the level recursion, used ONLY by the backward pass of
    ``inverse_blocks``) against torch autograd through the oracle's factor + selected_inverse: values and raw gradients."""
    from cyclic_gps._adjoint import selected_inverse
    gen = torch.Generator().manual_seed(0)
    for l in (1, 3, 8):
        G, Bm, LLT = orc.leg_params(l, seed=3)
        for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 33, 100):
            gaps = -torch.log(torch.rand(max(n - 1, 1), generator=gen, dtype=torch.float64)) + 0.05
            R, O = orc.leg_posterior_precision(gaps, G, Bm, LLT)
            if n == 1:
                R, O = R[:1].clone(), torch.zeros((0, l, l), dtype=torch.float64)
            Rr, Or = R.clone().requires_grad_(True), O.clone().requires_grad_(True)
            Sd0, So0 = orc.selected_inverse(orc.factor(Rr, Or))
            cd, co = torch.randn(Sd0.shape, generator=gen, dtype=torch.float64), torch.randn(So0.shape, generator=gen, dtype=torch.float64)
            ((Sd0 * cd).sum() + (So0 * co).sum()).backward()
            R2, O2 = R.clone().requires_grad_(True), O.clone().requires_grad_(True)
            Sd, So = selected_inverse(R2[None], O2[None])
            ((Sd[0] * cd).sum() + (So[0] * co).sum()).backward()
            assert_close(Sd[0], Sd0, 1e-11, f"Sigma_d l={l} n={n}")
            assert_close(R2.grad, Rr.grad, 1e-10, f"gR l={l} n={n}")
            if n > 1:
                assert_close(So[0], So0, 1e-11, f"Sigma_o l={l} n={n}")
                assert_close(O2.grad, Or.grad, 1e-10, f"gO l={l} n={n}")
