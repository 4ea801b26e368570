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
        pack = _engine.forward_sweep(self.R, self.O, self.x, keep_factors=True)
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
