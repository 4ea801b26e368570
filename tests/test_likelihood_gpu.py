"""GPU: LEGFamily.log_likelihood / compute_insample_posterior through the CUDA CR engine,
mirroring the reference's tests/test_likelihood.py (dense GP likelihood and Kalman filter as
independent answers) plus the golden values and reference-autograd parameter gradients."""
import numpy as np
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu


def _model_from_golden(g, p, d):
    from cyclic_gps.models import LEGFamily
    m = LEGFamily(rank=5, obs_dim=d, train=True, data_type=torch.float64)
    for name in ("N_params", "R_params", "Lambda_params", "B"):
        getattr(m, name).data = torch.from_numpy(g[p + name].copy())
    m.register_model_matrices_from_params()
    return m


def test_log_likelihood_matches_reference_values_and_gradients(golden):
    g = golden["leg_model"]
    for p in g["cases"]:
        p = str(p)
        d = int(p.split("_d")[1][0])
        m = _model_from_golden(g, p, d)
        ts, xs = torch.from_numpy(g[p + "ts"]), torch.from_numpy(g[p + "xs"])
        ll = m.log_likelihood(ts=ts, xs=xs)
        assert ll.dim() == 0
        assert_close(ll, g[p + "ll"], 1e-10, p + "loglik")
        assert_close(ll, g[p + "ll_dense"].reshape(()), 1e-9, p + "loglik vs dense")
        ll.backward()
        for name in ("N_params", "R_params", "Lambda_params", "B"):
            assert_close(getattr(m, name).grad, g[p + "grad_" + name], 1e-8, p + "grad " + name)
        with torch.no_grad():
            mean, cov = m.compute_insample_posterior(ts, xs)
        assert_close(mean, g[p + "post_mean"], 1e-9, p + "posterior mean")
        assert_close(cov["Rs"], g[p + "post_cov_R"], 1e-9, p + "posterior cov diag")
        assert_close(cov["Os"], g[p + "post_cov_O"], 1e-9, p + "posterior cov off")


def test_log_marginal_likelihood_like_the_reference_test():
    # reference tests/test_likelihood.py:9-29 (fewer sizes to keep the dense oracle cheap)
    from cyclic_gps.data_utils import generate_data
    from cyclic_gps.kalman import init_kalman_filter, kf_log_marginal_likelihood
    from cyclic_gps.model_utils import compute_log_marginal_likelihood
    from cyclic_gps.models import LEGFamily
    torch.manual_seed(0)
    for spacing in ["regular", "irregular"]:
        for n in [10, 33, 50, 100]:
            for d in [1, 2, 3]:
                ts, xs = generate_data(num_datapoints=n, data_dim=d, data_type=torch.double, spacing=spacing)
                model = LEGFamily(rank=5, obs_dim=xs.shape[-1], train=False, data_type=torch.double)
                model.double()
                naive = compute_log_marginal_likelihood(N=model.N, R=model.R, B=model.B,
                                                        Lambda=model.calc_Lambda_Lambda_T(model.Lambda), ts=ts, xs=xs)
                leg = model.log_likelihood(ts=ts.double(), xs=xs)
                if spacing == "regular":
                    kf = init_kalman_filter(leg_model=model, use_approximation=False)
                    kf_ll = kf_log_marginal_likelihood(kf, xs)
                    assert torch.allclose(leg, torch.from_numpy(np.array(kf_ll)))
                assert torch.allclose(leg, naive.reshape(()))
