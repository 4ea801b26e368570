"""Model-side helpers with the reference's names (cyclic_gps/model_utils.py): the dense
O((n d)^3) log-marginal-likelihood used as a test oracle by the reference's
tests/test_likelihood.py, and small matrix utilities.  Plain torch on the caller's device;
none of this is on the CR hot path."""
import math

import torch


def compute_G(N, R):
    """G = N N^T + R - R^T + 1e-5 I   (reference model_utils.py:5-10, models.py:152-159)."""
    return N @ N.T + R - R.T + 1e-5 * torch.eye(N.shape[0], dtype=N.dtype, device=N.device)


def compute_eG(G_val, G_vec, G_vec_inv, diffs):
    """exp(-d/2 G) for many gaps d from one eigendecomposition G = V diag(val) V^{-1}
    (reference model_utils.py:12-29).  Returns (m, l, l), real part."""
    scale = torch.exp(-0.5 * diffs.reshape(-1, 1, 1) * G_val.reshape(1, 1, -1))
    return torch.real((G_vec.unsqueeze(0) * scale) @ G_vec_inv.unsqueeze(0))


def build_2x2_block(a, b, c, d):
    return torch.cat([torch.cat([a, b], dim=-1), torch.cat([c, d], dim=-1)], dim=-2)


def build_3x3_block(a, b, c, d, e, f, g, h, i):
    return torch.cat([torch.cat([a, b, c], dim=-1), torch.cat([d, e, f], dim=-1), torch.cat([g, h, i], dim=-1)], dim=-2)


def gaussian_stitch(joint_mean, joint_cov, marginal_mean, marginal_cov):
    """Mean and covariance of y under q(x, y) = p2(x) p1(y | x), with p1 = N(joint_mean, joint_cov) over (x, y) and
    p2 = N(marginal_mean, marginal_cov) over x (reference model_utils.py:64-107; the reference's formula, which
    assumes a zero joint mean for the x part).  Batched over leading axes."""
    m = marginal_cov.shape[-1]
    T = joint_cov[..., m:, :m] @ torch.linalg.inv(joint_cov[..., :m, :m])
    mean = joint_mean[..., m:] + (T @ marginal_mean.unsqueeze(-1)).squeeze(-1)
    cond = joint_cov[..., m:, m:] - T @ joint_cov[..., :m, m:]
    return mean, cond + T @ marginal_cov @ T.transpose(-1, -2)


def compute_prior_covariance(ts, G):
    """Dense stationary LEG prior covariance, block (i,j) = exp(-|t_i - t_j|/2 G) for i > j and its
    transpose above the diagonal (reference model_utils.py:110-128, vectorised over all pairs)."""
    n, l = len(ts), G.shape[0]
    lag = ts.reshape(-1, 1) - ts.reshape(1, -1)                        # (n, n), t_i - t_j
    low = torch.matrix_exp(-0.5 * G.reshape(1, 1, l, l) * lag.abs().reshape(n, n, 1, 1))
    blocks = torch.where((lag >= 0).reshape(n, n, 1, 1), low, low.transpose(-1, -2))
    return blocks.permute(0, 2, 1, 3).reshape(n * l, n * l)


def compute_log_marginal_likelihood(N, R, B, Lambda, ts, xs):
    """log N(vec(xs); 0, B~ Sigma B~^T + Lambda~) by dense linear algebra
    (reference model_utils.py:131-142).  `Lambda` is the observation-noise COVARIANCE."""
    n = len(xs)
    G = compute_G(N, R)
    Bt = torch.block_diag(*([B] * n))
    cov = Bt @ compute_prior_covariance(ts=ts, G=G) @ Bt.T + torch.block_diag(*([Lambda] * n))
    x = xs.reshape(-1, 1)
    quad = x.T @ torch.linalg.solve(cov, x)
    return -0.5 * quad - 0.5 * torch.logdet(2 * math.pi * cov)
