"""LEG model glue with the reference's class and method names (cyclic_gps/models.py), kept
to what calls the CR hot path: parameters -> G -> block-tridiagonal precision (Rs, Os) ->
``decompose`` / ``det`` / ``mahal_and_det`` / ``solve`` / ``inverse_blocks`` from
``cyclic_gps.cyclic_reduction`` (the B200 engine).  The prediction helpers of the reference
(forecast / interpolate / intercast / predictive_posterior / make_predictions, models.py:394-546) are plain torch
glue behind the in-sample posterior, vectorised over the targets (the reference loops over them in Python).
``pytorch_lightning`` is optional: without it the class is a plain
``torch.nn.Module`` with the same training_step / configure_optimizers hooks."""
import math

import torch
from torch.optim import LBFGS, Adam
from torch.optim.lr_scheduler import ReduceLROnPlateau

from cyclic_gps.cyclic_reduction import decompose, det, inverse_blocks, mahal_and_det, solve, solve_and_inverse_blocks
from cyclic_gps.model_utils import build_2x2_block, build_3x3_block, compute_eG, gaussian_stitch
from cyclic_gps.peg import peg_precision

# log_likelihood: take log det of the prior precision from the precision builder (one cyclic reduction per evaluation) instead of
# a second factorisation as the reference does (models.py:349-353).  Same value (an algebraic identity, see peg.peg_precision).
FUSED_PRIOR_LOGDET = True


def _rows_times(x, W):
    """x (..., n, d) @ W (d, k).  With a single observation channel (d = 1, the shape of the reference's timing scripts) the
    product is a broadcast multiply: a GEMM with inner dimension 1 runs far below the memory roofline."""
    return x * W.reshape(W.shape[-1]) if x.shape[-1] == 1 else x @ W

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    class _Base(torch.nn.Module):
        def log(self, *args, **kwargs):
            return None


class LEGFamily(_Base):
    r"""z ~ PEG(N, R),  x(t) ~ N(B z(t), Lambda Lambda^T)   (reference models.py:20-27)."""

    def __init__(self, rank: int, obs_dim: int, prior_process_noise_level: float = 1.0, prior_length_scale: float = 0.2,
                 train: bool = False, optimizer: str = "ADAM", data_type=torch.float32, lr=1e-2) -> None:
        super().__init__()
        self.rank, self.obs_dim = rank, obs_dim
        self.prior_process_noise_level = prior_process_noise_level
        self.prior_length_scale = prior_length_scale
        self.optimizer, self.data_type, self.lr = optimizer, data_type, lr
        tri = lambda n, off: self.inds_to_tuple(torch.tril_indices(row=n, col=n, offset=off))
        self.N_idxs, self.R_idxs, self.Lambda_idxs = tri(rank, 0), tri(rank, -1), tri(obs_dim, 0)
        par = lambda k: torch.nn.Parameter(torch.zeros(k, dtype=data_type), requires_grad=train)
        self.N_params, self.R_params = par(len(self.N_idxs[0])), par(len(self.R_idxs[0]))
        self.Lambda_params = par(len(self.Lambda_idxs[0]))
        self.B = torch.nn.Parameter(torch.zeros((obs_dim, rank), dtype=data_type), requires_grad=train)
        self.get_initial_guess()
        self.register_model_matrices_from_params()

    @staticmethod
    def inds_to_tuple(raw_inds):
        return (raw_inds[0], raw_inds[1])

    # ---- initial values (reference models.py:81-121)
    def get_initial_guess(self):
        self.set_initial_N()
        self.set_initial_R()
        self.set_initial_B()
        self.set_initial_Lambda()

    def set_initial_N(self):
        N = torch.eye(self.rank, dtype=self.data_type) * self.prior_process_noise_level
        self.N_params.data = torch.linalg.cholesky(N @ N.T)[self.N_idxs]

    def set_initial_R(self):
        A = torch.randn((self.rank, self.rank), dtype=self.data_type)
        self.R_params.data = ((A - A.T) * self.prior_length_scale)[self.R_idxs]

    def set_initial_Lambda(self):
        Lam = 0.1 * torch.eye(self.obs_dim, dtype=self.data_type)
        self.Lambda_params.data = torch.linalg.cholesky(Lam @ Lam.T)[self.Lambda_idxs]

    def set_initial_B(self):
        ones = torch.ones((self.obs_dim, self.rank), dtype=self.data_type)
        self.B.data = 0.5 * ones / torch.sqrt(torch.sum(ones ** 2, dim=1, keepdim=True))

    @property
    def parameter_count(self) -> int:
        return len(self.N_params) + len(self.R_params) + len(self.Lambda_params) + torch.numel(self.B)

    # ---- parameters -> matrices (reference models.py:135-178)
    def _scatter(self, n, idxs, values):
        M = torch.zeros(n, n, dtype=values.dtype, device=values.device)
        M[idxs] = values
        return M

    def N_from_params(self):
        self.register_buffer("N", self._scatter(self.rank, self.N_idxs, self.N_params))

    def R_from_params(self):
        self.register_buffer("R", self._scatter(self.rank, self.R_idxs, self.R_params))

    def Lambda_from_params(self):
        self.register_buffer("Lambda", self._scatter(self.obs_dim, self.Lambda_idxs, torch.nn.functional.softplus(self.Lambda_params)))

    def calc_G(self):
        eye = torch.eye(self.rank, dtype=self.N.dtype, device=self.N.device)
        self.register_buffer("G", self.N @ self.N.T + self.R - self.R.T + 1e-5 * eye)

    @staticmethod
    def calc_Lambda_Lambda_T(Lambda):
        if Lambda.dim() == 2:
            return Lambda @ Lambda.T + 1e-9 * torch.eye(Lambda.shape[0], dtype=Lambda.dtype, device=Lambda.device)
        return torch.diag(Lambda ** 2 + 1e-9)

    def register_model_matrices_from_params(self):
        self.Lambda_from_params()
        self.N_from_params()
        self.R_from_params()
        self.calc_G()

    # ---- precision blocks (reference models.py:181-239, 254-280)
    @staticmethod
    def _compute_device(t):
        """Where the per-row work runs: the current CUDA device when there is one (inputs may live on the CPU, as in the
        reference's tests; only ts / xs cross PCIe), else the tensor's own device (model glue only -- the CR calls raise)."""
        if t.is_cuda or not torch.cuda.is_available():
            return t.device
        return torch.device("cuda", torch.cuda.current_device())

    def _precision_blocks(self, ts, shift=None, logdet=False, out=None):
        """(Rs, Os) on the compute device, built by ONE kernel from the time gaps (cyclic_gps.peg) with a
        hand-written backward to G and the diagonal shift.  ts (n,) or (B,n).  logdet=True: also log det of the unshifted
        (prior) precision, a by-product of the same kernel."""
        dev = self._compute_device(ts)
        t = ts.to(dev)
        gaps = (t[..., 1:] - t[..., :-1]).to(self.G.dtype)
        return peg_precision(gaps, self.G, shift, logdet=logdet, out=out)

    def compute_PEG_precision(self, ts):
        """Diagonal (n,l,l) and lower off-diagonal (n-1,l,l) blocks of the PEG precision, on the caller's device; a leading
        batch axis of independent series, ts (B,n), gives (B,n,l,l) and (B,n-1,l,l) (the reference is batch-of-one,
        SURVEY 8(f2))."""
        Rs, Os = self._precision_blocks(ts)
        return Rs.to(ts.device), Os.to(ts.device)

    def _obs_terms(self):
        LLT = self.calc_Lambda_Lambda_T(self.Lambda)
        return LLT, self.B.T @ torch.linalg.solve(LLT, self.B)

    def _obs_factor(self):
        """(LL^T)^{-1}, W = (LL^T)^{-1} B, shift = B^T W and log det(2 pi LL^T) from ONE Cholesky of the d x d matrix, without the
        device -> host status reads of three LU-based calls (solve / inv / logdet): at small n those calls are the step's cost."""
        LLT = self.calc_Lambda_Lambda_T(self.Lambda)
        C, _ = torch.linalg.cholesky_ex(LLT)
        inv = torch.cholesky_inverse(C)
        W = inv @ self.B
        logdet = LLT.shape[0] * math.log(2 * math.pi) + 2 * torch.log(torch.diagonal(C)).sum()
        return inv, W, self.B.T @ W, logdet

    def compute_posterior_precision(self, ts):
        _, shift = self._obs_terms()
        Rs, Os = self._precision_blocks(ts, shift)
        return Rs.to(ts.device), Os.to(ts.device)

    def compute_v(self, xs):
        """v = B^T (LL^T)^{-1} x for every row (reference models.py:270-280): the (d x l) map is formed once on the
        parameters' device, the per-row work is one matmul on the device of xs."""
        LLT = self.calc_Lambda_Lambda_T(self.Lambda)
        return _rows_times(xs, torch.linalg.solve(LLT, self.B).to(xs.device))

    def compute_insample_posterior(self, ts, xs):
        """Posterior mean (n,l) and {"Rs","Os"} blocks of the posterior covariance
        (reference models.py:282-298); computed on the GPU, returned on the caller's device."""
        dev = self._compute_device(ts)
        _, shift = self._obs_terms()
        Rs, Os = self._precision_blocks(ts, shift)
        mean, Sd, So = solve_and_inverse_blocks(Rs, Os, self.compute_v(xs.to(dev)))     # two sweeps: factor + half-solve, back-solve + selected inverse
        out = ts.device
        return mean.to(out), {"Rs": Sd.to(out), "Os": So.to(out)}

    def log_likelihood(self, ts, xs):
        """log p(xs | ts) (reference models.py:301-372: two CR factorisations; here the prior's log-determinant comes out of the
        precision builder and only the posterior precision is factorised -- FUSED_PRIOR_LOGDET = False restores the two sweeps).  ts (n,), xs (n,d) give a
        scalar as in the reference; a batch of independent series, ts (B,n) and xs (B,n,d), gives (B,) values from
        ONE batched pass of the CR engine (the series share the model parameters).  Only ts and xs travel to the GPU:
        the precision blocks are built there (``_precision_blocks``)."""
        self.register_model_matrices_from_params()
        dev = self._compute_device(ts)
        LLT_inv, W, shift, logdet_2pi_LLT = self._obs_factor()
        xs_d = xs.to(dev)
        white = _rows_times(xs_d, LLT_inv.to(dev))            # x^T (LL^T)^{-1} per row (LL^T is d x d and symmetric)
        obs_mahal = torch.sum(white * xs_d, dim=(-1, -2))
        obs_logdet = (logdet_2pi_LLT * xs.shape[-2]).to(dev)
        v = _rows_times(xs_d, W.to(dev))
        if FUSED_PRIOR_LOGDET:
            # ONE pass over the gaps gives the posterior precision K = Sigma^{-1} + shift AND log det Sigma^{-1} (the reference
            # factorises both matrices, models.py:349-367): one cyclic reduction instead of two, forward and backward
            Rs, Os, prior_logdet = self._precision_blocks(ts, shift, logdet=True)
            K_mahal, K_logdet = mahal_and_det(Rs=Rs, Os=Os, x=v)
            prior_logdet = prior_logdet.to(K_logdet.dtype)
        else:
            Rs, Os = self._precision_blocks(ts)
            prior_logdet = det(decompose(Rs, Os))
            K_mahal, K_logdet = mahal_and_det(Rs=Rs + shift.to(dev), Os=Os, x=v)
        return (-0.5 * ((obs_mahal - K_mahal) + (obs_logdet + K_logdet - prior_logdet))).to(ts.device)

    # ---- predictions at new times (reference models.py:394-546), vectorised over the targets
    def forecast(self, eG, ip_mean, ip_cov):
        """Latent mean / covariance one gap away from an in-sample time; eG = exp(-gap/2 G) (transposed by the
        caller when looking backwards), reference models.py:394-407."""
        eye = torch.eye(self.rank, dtype=eG.dtype, device=eG.device).expand_as(eG)
        joint_cov = build_2x2_block(eye, eG.transpose(-1, -2), eG, eye)
        joint_mean = torch.zeros(eG.shape[:-2] + (2 * self.rank,), dtype=eG.dtype, device=eG.device)
        return gaussian_stitch(joint_mean, joint_cov, ip_mean, ip_cov)

    def interpolate(self, eG1, eG2, prev_ip_mean, prev_ip_cov_diag, prev_ip_cov_offdiag, next_ip_mean, next_ip_cov_diag):
        """Latent mean / covariance between two in-sample times (reference models.py:409-451)."""
        eye = torch.eye(self.rank, dtype=eG1.dtype, device=eG1.device).expand_as(eG1)
        eG3 = eG1 @ eG2
        t = lambda a: a.transpose(-1, -2)
        joint_cov = build_3x3_block(eye, t(eG3), t(eG1), eG3, eye, eG2, eG1, t(eG2), eye)
        joint_mean = torch.zeros(eG1.shape[:-2] + (3 * self.rank,), dtype=eG1.dtype, device=eG1.device)
        ip_mean = torch.cat([prev_ip_mean, next_ip_mean], dim=-1)
        ip_cov = build_2x2_block(prev_ip_cov_diag, t(prev_ip_cov_offdiag), prev_ip_cov_offdiag, next_ip_cov_diag)
        return gaussian_stitch(joint_mean, joint_cov, ip_mean, ip_cov)

    def intercast(self, ip_mean, ip_cov, ts, target_ts, thresh=1e-10):
        """Posterior of the latent process at `target_ts` (sorted) from the in-sample posterior
        (reference models.py:454-515: forecast before / after the data, interpolate inside, exact hits at the ends)."""
        assert bool((target_ts[1:] - target_ts[:-1] > 0).all())
        dev, dt = ip_mean.device, ip_mean.dtype
        ts, target_ts = ts.to(dev), target_ts.to(dev)
        G_val, G_vec = torch.linalg.eig(self.G.detach().cpu())         # G is tiny; one eigendecomposition on the host
        G_vec_inv = torch.linalg.inv(G_vec)
        G_val, G_vec, G_vec_inv = G_val.to(dev), G_vec.to(dev), G_vec_inv.to(dev)
        eG = lambda gaps: compute_eG(G_val, G_vec, G_vec_inv, gaps).to(dt)
        n = ts.shape[0]
        Rs, Os = ip_cov["Rs"], ip_cov["Os"]
        idx = torch.searchsorted(ts, target_ts)
        close = lambda a, b: (a - b).abs() <= 1e-8 + 1e-5 * b.abs()     # torch.allclose defaults
        before, after = idx == 0, idx == n
        at_first, at_last = before & close(target_ts, ts[0]), close(target_ts, ts[-1])
        # forecasting backwards from the first / forwards from the last in-sample time
        mb, vb = self.forecast(eG((ts[0] - target_ts).clamp(min=0)).transpose(-1, -2), ip_mean[0], Rs[0])
        ma, va = self.forecast(eG((target_ts - ts[-1]).clamp(min=0)), ip_mean[-1], Rs[-1])
        # interpolation between ts[idx-1] and ts[idx]
        lo, hi = (idx - 1).clamp(0, n - 2), idx.clamp(1, n - 1)
        mi, vi = self.interpolate(eG((target_ts - ts[lo]).clamp(min=0)), eG((ts[hi] - target_ts).clamp(min=0)),
                                  ip_mean[lo], Rs[lo], Os[lo], ip_mean[hi], Rs[hi])
        pick = lambda c, a, b: torch.where(c.reshape((-1,) + (1,) * (a.dim() - 1)), a, b)
        mean, cov = pick(before, mb, pick(after, ma, mi)), pick(before, vb, pick(after, va, vi))
        mean, cov = pick(at_first, ip_mean[0].expand_as(mean), mean), pick(at_first, Rs[0].expand_as(cov), cov)
        exact_last = at_last & ~before
        mean, cov = pick(exact_last, ip_mean[-1].expand_as(mean), mean), pick(exact_last, Rs[-1].expand_as(cov), cov)
        return mean, cov

    def predictive_posterior(self, ts, xs, target_ts):
        """E[z(t) | xs], Cov[z(t) | xs] at the target times (reference models.py:517-530)."""
        mean, cov = self.compute_insample_posterior(ts, xs)
        return self.intercast(mean, cov, ts, target_ts)

    def make_predictions(self, ts, xs, target_ts):
        """Predicted observations: means (m, obs_dim) and covariances (m, obs_dim, obs_dim) at the target times,
        without the observation noise (reference models.py:532-546)."""
        self.calc_G()
        zm, zv = self.predictive_posterior(ts, xs, target_ts)
        B = self.B.to(zm.device)
        return zm @ B.T, B.unsqueeze(0) @ zv @ B.T.unsqueeze(0)

    # ---- training hooks (reference models.py:374-392)
    # Opt-in: evaluate the training loss through cyclic_gps.graphs.GraphedLogLikelihood (one CUDA-graph replay for the device part of
    # the likelihood and its backward).  The reference's training loop feeds the SAME (ts, xs) every step (models.py:374-392); the
    # runner is rebuilt whenever the batch differs from the one it was captured for.
    graphed_training = False

    def _graphed_loss(self, t, x):
        from cyclic_gps.graphs import GraphedLogLikelihood
        cache = getattr(self, "_graph_cache", None)
        if cache is None or cache[0].shape != t.shape or cache[1].shape != x.shape or not (torch.equal(cache[0], t) and torch.equal(cache[1], x)):
            cache = (t.clone(), x.clone(), GraphedLogLikelihood(self, t, x))
            object.__setattr__(self, "_graph_cache", cache)
        return cache[2]()

    def training_step(self, train_batch, batch_idx):
        t, x = train_batch
        nobs = x.shape[0] * x.shape[1] * x.shape[2]
        if self.graphed_training and torch.cuda.is_available():
            loss = -self._graphed_loss(t.squeeze(0), x.squeeze(0)) / nobs
        else:
            loss = -self.log_likelihood(t.squeeze(0), x.squeeze(0)) / nobs
        self.log("NLL", loss)
        return loss

    def configure_optimizers(self):
        params = [p for p in self.parameters() if p.requires_grad]
        opt = Adam(params, lr=self.lr) if self.optimizer == "ADAM" else LBFGS(params, lr=self.lr, max_iter=20)
        return {"optimizer": opt, "lr_scheduler": ReduceLROnPlateau(opt, "min"), "monitor": "NLL"}
