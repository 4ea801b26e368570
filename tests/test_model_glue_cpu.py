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
