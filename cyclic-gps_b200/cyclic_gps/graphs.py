"""CUDA-graph replay of the fused likelihood pass for fixed shapes (small problems are bound by launch and
host overhead, not by the GPU).

``GraphedMahalAndDet(R, O, x)`` captures ONE forward sweep (crb200_sweep_fwd: factor, half-solve, log-det, Mahalanobis
term) and ONE backward sweep (crb200_sweep_bwd: back-solve, selected inverse, gradient assembly) of
``cyclic_reduction.mahal_and_det`` into a CUDA graph.  A call copies the new inputs into the graph's static buffers and
replays it: no Python per level, no autograd bookkeeping, no allocator traffic, one graph launch per step.  It returns
``(mahal, logdet, gR, gO, gx)`` for the cotangents given at construction (default 1, 1) -- i.e. the value and the
gradient of ``g_m * mahal + g_d * logdet`` -- as views of static buffers that the next call overwrites.

The positive-definiteness report cannot be read inside a graph (it needs a device->host copy); ``check()`` reads the
report of the last replay on demand."""
import torch

from . import _engine


class GraphedMahalAndDet:
    def __init__(self, Rs, Os, x, g_mahal=1.0, g_det=1.0, warmup=3):
        dev = _engine.require_cuda()
        batched = Rs.dim() == 4
        self._batched = batched
        b = (lambda t: t) if batched else (lambda t: t.unsqueeze(0))
        self.R = b(Rs.detach().to(dev)).contiguous().clone()
        self.O = b(Os.detach().to(dev)).contiguous().clone()
        self.x = b(x.detach().to(dev, Rs.dtype)).contiguous().clone()
        B = self.R.shape[0]
        self.gm = torch.full((B,), float(g_mahal), dtype=torch.float64, device=dev)
        self.gd = torch.full((B,), float(g_det), dtype=torch.float64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # sets the kernels' attributes and warms the allocator outside the capture
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._out = self._run()

    def _run(self):
        pack = _engine.forward_sweep(self.R, self.O, self.x, keep_factors=True, internal=True)
        gR, gO, gx = _engine.backward_sweep(pack, sigma=True, w=True, grad=(self.gm, self.gd))
        return pack.mahal.to(self.R.dtype), pack.logdet.to(self.R.dtype), gR, gO, gx, pack

    def __call__(self, Rs, Os, x):
        b = (lambda t: t) if self._batched else (lambda t: t.unsqueeze(0))
        if Rs.data_ptr() != self.R.data_ptr():
            self.R.copy_(b(Rs), non_blocking=True)
        if Os.data_ptr() != self.O.data_ptr():
            self.O.copy_(b(Os), non_blocking=True)
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(b(x), non_blocking=True)
        self.graph.replay()
        mh, ld, gR, gO, gx, _ = self._out
        if self._batched:
            return mh, ld, gR, gO, gx
        return mh[0], ld[0], gR[0], gO[0], gx[0]

    def check(self):
        """Raise NotPositiveDefiniteError if the last replay met a non-positive-definite block (one device->host read)."""
        self._out[5].check()


class _GraphedLLFn(torch.autograd.Function):
    """Host-side autograd node around one replay: (G, shift, W, LLT^{-1}) -> sum_b of the parameter-dependent device part of the
    log-likelihood; the replay has already produced the four cotangents."""

    @staticmethod
    def forward(ctx, G, shift, W, Linv, runner):
        val, grads = runner._replay(G, shift, W, Linv)
        ctx.grads = tuple(g.to(t.device, t.dtype) for g, t in zip(grads, (G, shift, W, Linv)))
        return val.to(G.device, G.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        gG, gs, gW, gL = ctx.grads
        return g * gG, g * gs, g * gW, g * gL, None


class GraphedLogLikelihood:
    """``sum_b LEGFamily.log_likelihood(ts, xs)[b]`` and its gradient wrt the model parameters with ONE CUDA-graph replay per
    evaluation, for fixed time stamps and observations (the training loop of the reference: models.py:374-392 evaluates the same
    data every step).  Small problems are bound by launches, host synchronisations and autograd bookkeeping of ~150 eager ops
    (DESIGN section 6): here the device part -- precision builder forward (+ prior log-determinant), observation maps, ONE
    cyclic-reduction forward + backward sweep, builder backward with its complex finish, the reductions for the cotangents of
    the observation maps, and the copy of everything the host needs into pinned memory -- is captured once; per call the host
    only rebuilds the l x l / d x d parameter algebra (autograd-tracked torch ops on the model's device), eigendecomposes G
    (numpy), refreshes four small device arrays and replays.

        runner = GraphedLogLikelihood(model, ts, xs)
        loss = -runner() / n; loss.backward()        # gradients land in model.parameters()

    The captured constants depend on the STRUCTURE of G's spectrum (how many complex pairs); if that changes between calls the
    graph is captured again.  A non-positive-definite block raises NotPositiveDefiniteError from the call."""

    def __init__(self, model, ts, xs, warmup=3):
        from . import peg
        self._peg = peg
        dev = _engine.require_cuda()
        self.model, self.dev = model, dev
        dtype = model.G.dtype
        t = (ts if ts.dim() == 2 else ts.unsqueeze(0)).to(dev)
        x = (xs if xs.dim() == 3 else xs.unsqueeze(0)).to(dev, dtype)
        self.B, self.n = t.shape
        self.l, self.d, self.dtype = model.rank, model.obs_dim, dtype
        if not peg.device_builder_available(self.l):
            raise ValueError(f"no device precision builder for rank {self.l}")
        self.gaps = (t[:, 1:] - t[:, :-1]).to(dtype).contiguous()
        self.xs = x.contiguous()
        l, d, B = self.l, self.d, self.B
        self.p_shift = torch.zeros((l, l), dtype=torch.float64, device=dev)
        self.p_W = torch.zeros((d, l), dtype=dtype, device=dev)
        self.p_Linv = torch.zeros((d, d), dtype=dtype, device=dev)
        self.gm = torch.full((B,), 0.5, dtype=torch.float64, device=dev)         # d sum(ll) / d K_mahal
        self.gd = torch.full((B,), -0.5, dtype=torch.float64, device=dev)        # d sum(ll) / d K_logdet
        self.gpl = torch.full((B,), 0.5, dtype=torch.float64, device=dev)        # d sum(ll) / d prior_logdet
        # d sum(ll) / d LLT^{-1} = -1/2 sum x x^T does not depend on the parameters
        self.g_Linv = -0.5 * torch.einsum("bnd,bne->de", x.double(), x.double())
        self.nout = 2 + 2 * l * l + d * l
        self.out_host = torch.zeros(self.nout, dtype=torch.float64).pin_memory()
        self.consts = None
        self.graph = None
        self._warmup = warmup

    # -- device part (captured)
    def _run(self):
        peg, c, l, d, dtype = self._peg, self.consts, self.l, self.d, self.dtype
        R, O, prior_ld, binfo = peg.builder_forward(c, self.gaps, self.p_shift, dtype, True)
        one = self.d == 1
        v = self.xs * self.p_W.reshape(l) if one else self.xs @ self.p_W
        white = self.xs * self.p_Linv.reshape(1) if one else self.xs @ self.p_Linv
        obs_mahal = (white * self.xs).sum(dim=(-1, -2), dtype=torch.float64)
        pack = _engine.forward_sweep(R, O, v, keep_factors=True, internal=True)
        gR, gO, gv = _engine.backward_sweep(pack, sigma=True, w=True, grad=(self.gm, self.gd))
        gG, gshift = peg.builder_backward(c, self.gaps, O, gR, gO, self.gpl, dtype)
        gW = torch.einsum("bnd,bnl->dl", self.xs.double(), gv.double())
        ll = (-0.5 * obs_mahal + 0.5 * pack.mahal - 0.5 * pack.logdet + 0.5 * prior_ld).sum()
        bad = (binfo != 0).sum().double() + (pack.info != 0).sum().double() if pack.info is not None else (binfo != 0).sum().double()
        out = torch.cat([ll.reshape(1), bad.reshape(1), gG.reshape(-1), gshift.reshape(-1), gW.reshape(-1)])
        self.out_host.copy_(out, non_blocking=True)
        return out, pack

    def _capture(self, G_host):
        self.consts = self._peg._EigConsts(None, self.dev, G_host)
        if not self.consts.folded or self.consts.cond > self._peg.EIG_COND_MAX:
            raise ValueError("the eigenbasis of G is too ill-conditioned for the device precision builder")
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._run()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._keep = self._run()

    def _replay(self, G, shift, W, Linv):
        G_host = G.detach().to("cpu", torch.float64)
        with torch.no_grad():
            self.p_shift.copy_(shift.detach(), non_blocking=True)
            self.p_W.copy_(W.detach(), non_blocking=True)
            self.p_Linv.copy_(Linv.detach(), non_blocking=True)
        if self.graph is None or not self.consts.refresh(G_host):
            self._capture(G_host)
        elif self.consts.cond > self._peg.EIG_COND_MAX:
            raise ValueError("the eigenbasis of G is too ill-conditioned for the device precision builder")
        self.graph.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        o = self.out_host
        if float(o[1]) != 0.0:
            raise _engine.NotPositiveDefiniteError("log-likelihood: a diagonal block (or I - A A^T of a gap) was not positive definite")
        l, d = self.l, self.d
        val = o[0].clone()
        gG = o[2:2 + l * l].reshape(l, l).clone()
        gs = o[2 + l * l:2 + 2 * l * l].reshape(l, l).clone()
        gW = o[2 + 2 * l * l:].reshape(d, l).clone()
        return val, (gG, gs, gW, self.g_Linv)

    def __call__(self):
        """sum over the series of the log-likelihood, attached to the autograd graph of the model parameters."""
        import math
        m = self.model
        m.register_model_matrices_from_params()
        Linv, W, shift, logdet_2pi_LLT = m._obs_factor()
        part = _GraphedLLFn.apply(m.G, shift, W, Linv, self)
        return part - 0.5 * self.B * self.n * logdet_2pi_LLT.to(part.device, part.dtype)
