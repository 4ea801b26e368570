"""Kalman-filter comparator with the reference's entry points (cyclic_gps/kalman.py:7-60).
The reference delegates to filterpy (absent here); this is a small numpy filter with the same
predict / update / log-likelihood semantics (P0 = I, x0 = 0).  Comparison baseline only."""
import math

import numpy as np
from scipy.linalg import expm


class KalmanFilter:
    def __init__(self, dim_x, dim_z):
        self.dim_x, self.dim_z = dim_x, dim_z
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)
        self.log_likelihood = 0.0

    def predict(self):
        self.x = self.F @ self.x
        self.P = self.F @ self.P @ self.F.T + self.Q

    def update(self, z):
        z = np.asarray(z, dtype=float).reshape(self.dim_z, 1)
        resid = z - self.H @ self.x
        S = self.H @ self.P @ self.H.T + self.R
        gain = np.linalg.solve(S, self.H @ self.P).T
        self.x = self.x + gain @ resid
        J = np.eye(self.dim_x) - gain @ self.H
        self.P = J @ self.P @ J.T + gain @ self.R @ gain.T
        self.log_likelihood = float(-0.5 * (resid.T @ np.linalg.solve(S, resid)).item()
                                    - 0.5 * np.linalg.slogdet(S)[1] - 0.5 * self.dim_z * math.log(2 * math.pi))


def init_kalman_filter(leg_model, time_step=1, use_approximation=True):
    leg_model.register_model_matrices_from_params()
    G = leg_model.G.detach().cpu().numpy()
    kf = KalmanFilter(dim_x=leg_model.rank, dim_z=leg_model.obs_dim)
    if use_approximation:
        kf.F = np.eye(leg_model.rank) - 0.5 * time_step * G
        kf.Q = time_step * (leg_model.N @ leg_model.N.T).detach().cpu().numpy()
    else:
        kf.F = expm(-0.5 * time_step * G)
        kf.Q = np.eye(leg_model.rank) - kf.F @ kf.F.T
    kf.H = leg_model.B.detach().cpu().numpy()
    kf.R = leg_model.calc_Lambda_Lambda_T(leg_model.Lambda).detach().cpu().numpy()
    return kf


def kf_log_marginal_likelihood(kf, data):
    total = 0.0
    for i in range(data.shape[0]):
        kf.predict()
        kf.update(np.asarray(data[i]))
        total += kf.log_likelihood
    return total
