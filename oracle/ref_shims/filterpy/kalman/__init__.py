"""Stand-in for filterpy.kalman.KalmanFilter (oracle scaffolding): textbook predict/update."""
import math
import numpy as np


class KalmanFilter:
    def __init__(self, dim_x, dim_z, dim_u=0):
        self.dim_x, self.dim_z = dim_x, dim_z
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)
        self.log_likelihood = 0.0

    def predict(self):
        self.x = self.F @ self.x
        self.P = self.F @ self.P @ self.F.T + self.Q

    def update(self, z):
        z = np.asarray(z, dtype=float).reshape(self.dim_z, 1)
        y = z - self.H @ self.x
        S = self.H @ self.P @ self.H.T + self.R
        K = self.P @ self.H.T @ np.linalg.inv(S)
        self.x = self.x + K @ y
        IKH = np.eye(self.dim_x) - K @ self.H
        self.P = IKH @ self.P @ IKH.T + K @ self.R @ K.T
        sign, logdet = np.linalg.slogdet(S)
        self.log_likelihood = float(
            -0.5 * (y.T @ np.linalg.solve(S, y)).item() - 0.5 * logdet - 0.5 * self.dim_z * math.log(2 * math.pi)
        )
