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


def test_model_resident_on_the_gpu(golden):
    """The same model with parameters, time stamps and data on the device: no host round trips inside
    log_likelihood; values and parameter gradients must still match the reference."""
    g = golden["leg_model"]
    for p in ("irregular_n33_d2_", "regular_n33_d1_"):
        d = int(p.split("_d")[1][0])
        m = _model_from_golden(g, p, d).cuda()
        m.register_model_matrices_from_params()
        ts, xs = torch.from_numpy(g[p + "ts"]).cuda(), torch.from_numpy(g[p + "xs"]).cuda()
        ll = m.log_likelihood(ts=ts, xs=xs)
        assert ll.is_cuda
        assert_close(ll, g[p + "ll"], 1e-10, p + "loglik on device")
        ll.backward()
        for name in ("N_params", "R_params", "Lambda_params", "B"):
            assert_close(getattr(m, name).grad, g[p + "grad_" + name], 1e-8, p + "grad " + name)


def test_co2_shaped_training_step():
    """BASELINE config 3 shape: n = 502 with a 240-unit gap mid-series, rank 16, fp64, obs_dim 1: one training step
    (loss + backward through both CR factorisations) and the in-sample posterior, on the device, against the
    same model evaluated with the CPU oracle for the CR calls."""
    import math
    from cyclic_gps.models import LEGFamily
    from oracle import cr_oracle as orc
    torch.manual_seed(3)
    n, rank = 502, 16
    gaps = torch.ones(n - 1, dtype=torch.float64)
    gaps[261] = 240.0
    ts = torch.cat([torch.zeros(1, dtype=torch.float64), torch.cumsum(gaps, 0)]) / 12.0
    xs = torch.randn(n, 1, dtype=torch.float64)
    model = LEGFamily(rank=rank, obs_dim=1, train=True, data_type=torch.float64)
    loss = model.training_step((ts.unsqueeze(0), xs.unsqueeze(0)), 0)
    loss.backward()
    # oracle evaluation of the same likelihood (CPU, torch autograd through the oracle)
    ref = LEGFamily(rank=rank, obs_dim=1, train=True, data_type=torch.float64)
    ref.load_state_dict(model.state_dict())
    ref.register_model_matrices_from_params()
    LLT, shift = ref._obs_terms()
    white = torch.linalg.solve(LLT, xs.T).T
    Rs, Os = ref.compute_PEG_precision(ts)
    mh, ld = orc.mahal_and_logdet(Rs + shift.unsqueeze(0), Os, white @ ref.B)
    prior = orc.logdet(orc.factor(Rs, Os))
    ll = -0.5 * ((torch.sum(white * xs) - mh) + (torch.logdet(2 * math.pi * LLT) * n + ld - prior))
    (-ll / n).backward()
    assert_close(loss, -ll / n, 1e-10, "loss")
    for name in ("N_params", "R_params", "Lambda_params", "B"):
        assert_close(getattr(model, name).grad, getattr(ref, name).grad, 1e-8, "grad " + name)
    with torch.no_grad():
        mean, cov = model.compute_insample_posterior(ts, xs)
        dec = orc.factor(*ref.compute_posterior_precision(ts))
        assert_close(mean, orc.solve(dec, ref.compute_v(xs)), 1e-9, "posterior mean")
        sd, so = orc.selected_inverse(dec)
        assert_close(cov["Rs"], sd, 1e-9, "posterior cov diag")
        assert_close(cov["Os"], so, 1e-9, "posterior cov off")


def test_batched_log_likelihood_equals_loop():
    """LEGFamily.log_likelihood with a leading batch axis (ts (B,n), xs (B,n,d)): values and parameter gradients
    must equal the loop over single series (the reference's batch-of-one semantics)."""
    from cyclic_gps.models import LEGFamily
    torch.manual_seed(3)
    B, n, rank, d = 4, 300, 5, 2
    model = LEGFamily(rank=rank, obs_dim=d, train=True, data_type=torch.float64).cuda()
    ts = torch.cumsum(torch.rand(B, n, dtype=torch.float64, device="cuda") + 0.05, dim=1)
    xs = torch.randn(B, n, d, dtype=torch.float64, device="cuda")
    wts = torch.tensor([0.5, 1.0, 1.5, 2.0], dtype=torch.float64, device="cuda")
    ll_b = model.log_likelihood(ts=ts, xs=xs)
    assert ll_b.shape == (B,)
    (ll_b * wts).sum().backward()
    grads_b = [p.grad.clone() for p in model.parameters()]
    for p in model.parameters():
        p.grad = None
    total = 0
    for b in range(B):
        ll = model.log_likelihood(ts=ts[b], xs=xs[b])
        assert abs(float(ll.detach()) - float(ll_b[b].detach())) <= 1e-10 * abs(float(ll.detach()))
        total = total + wts[b] * ll
    total.backward()
    for g_b, p in zip(grads_b, model.parameters()):
        assert float((g_b - p.grad).abs().max()) <= 1e-9 * max(float(p.grad.abs().max()), 1e-30)


def test_predictions_match_the_reference(golden):
    """LEGFamily.predictive_posterior / make_predictions (vectorised over targets here, a Python loop in the
    reference, models.py:454-546) against golden vectors of the unmodified reference: targets before, at, between and
    beyond the observation times."""
    from cyclic_gps.models import LEGFamily
    g = golden["predictions"]
    for p in [str(c) for c in g["cases"]]:
        ts, xs, target = (torch.from_numpy(g[p + k]).cuda() for k in ("ts", "xs", "target"))
        rank = int(p.split("_r")[-1].rstrip("_"))
        m = LEGFamily(rank=rank, obs_dim=xs.shape[-1], train=False, data_type=torch.float64)
        for name in ("N_params", "R_params", "Lambda_params", "B"):
            getattr(m, name).data = torch.from_numpy(g[p + name])
        m.register_model_matrices_from_params()
        m = m.cuda()
        with torch.no_grad():
            zm, zv = m.predictive_posterior(ts, xs, target)
            xm, xv = m.make_predictions(ts, xs, target)
        for ours, key in ((zm, "z_mean"), (zv, "z_cov"), (xm, "x_mean"), (xv, "x_cov")):
            ref = torch.from_numpy(g[p + key])
            err = float((ours.cpu() - ref).abs().max() / ref.abs().max())
            assert err < 1e-8, (p, key, err)


@pytest.mark.parametrize("rank,d,B,n,dtype,tol_v,tol_g,on_gpu", [(5, 2, 1, 120, torch.float64, 1e-10, 1e-8, False), (8, 1, 3, 257, torch.float32, 1e-4, 2e-3, True),
                                                                 (16, 1, 1, 502, torch.float64, 1e-10, 1e-8, True), (3, 1, 2, 1000, torch.float64, 1e-10, 1e-8, True)])
def test_graphed_log_likelihood_matches_eager(rank, d, B, n, dtype, tol_v, tol_g, on_gpu):
    """cyclic_gps.graphs.GraphedLogLikelihood: one CUDA-graph replay per evaluation gives the value and the parameter gradients of
    the eager LEGFamily.log_likelihood(...).sum(), also after the parameters have moved (the constants of G are refreshed in place)."""
    from cyclic_gps.graphs import GraphedLogLikelihood
    from cyclic_gps.models import LEGFamily
    torch.manual_seed(rank + n)
    gaps = torch.rand((B, n - 1), dtype=torch.float64) + 0.05
    ts = torch.cat([torch.zeros((B, 1), dtype=torch.float64), torch.cumsum(gaps, 1)], 1)
    xs = torch.randn((B, n, d), dtype=dtype)
    model = LEGFamily(rank=rank, obs_dim=d, train=True, data_type=dtype)
    if on_gpu:
        model, ts, xs = model.cuda(), ts.cuda(), xs.cuda()
    runner = GraphedLogLikelihood(model, ts, xs)
    names = ("N_params", "R_params", "Lambda_params", "B")
    for it in range(3):
        model.zero_grad(set_to_none=True)
        want = model.log_likelihood(ts, xs).sum()
        (-want / n).backward()
        gw = [getattr(model, nm).grad.clone() for nm in names]
        model.zero_grad(set_to_none=True)
        got = runner()
        (-got / n).backward()
        assert_close(got, want, tol_v, f"graphed loglik, step {it}")
        for nm, g0 in zip(names, gw):
            assert_close(getattr(model, nm).grad, g0, tol_g, f"graphed grad {nm}, step {it}")
        with torch.no_grad():                             # move the parameters like an optimizer step would
            for nm in names:
                p = getattr(model, nm)
                p.add_(0.03 * torch.randn(p.shape, dtype=p.dtype, device=p.device, generator=None))


def test_graphed_training_step_is_the_eager_one():
    """LEGFamily.graphed_training = True: training_step through the CUDA-graph runner gives the eager loss and gradients, step after
    step with an optimiser moving the parameters, and notices a different batch."""
    from cyclic_gps.models import LEGFamily
    torch.manual_seed(11)
    n = 300
    ts = torch.cumsum(torch.rand(n, dtype=torch.float64) + 0.1, 0)
    xs = torch.randn(n, 2, dtype=torch.float64)
    a = LEGFamily(rank=6, obs_dim=2, train=True, data_type=torch.float64)
    b = LEGFamily(rank=6, obs_dim=2, train=True, data_type=torch.float64)
    b.load_state_dict(a.state_dict())
    b.graphed_training = True
    oa, ob = torch.optim.Adam(a.parameters(), lr=1e-2), torch.optim.Adam(b.parameters(), lr=1e-2)
    for it in range(4):
        if it == 3:
            xs = xs + 0.1                                   # a new batch: the runner must be rebuilt, not replayed
        batch = (ts.unsqueeze(0), xs.unsqueeze(0))
        oa.zero_grad(); ob.zero_grad()
        la, lb = a.training_step(batch, it), b.training_step(batch, it)
        la.backward(); lb.backward()
        assert_close(lb, la, 1e-10, f"loss, step {it}")
        for (na, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert_close(pb.grad, pa.grad, 1e-8, f"grad {na}, step {it}")
        oa.step(); ob.step()
