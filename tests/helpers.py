"""Shared helpers for the parity tests."""
import numpy as np
import torch

# Tolerances of BASELINE.json north_star: relative 1e-10 in fp64, 1e-4 in fp32, where
# "relative" is max-abs-error / max-abs-reference per tensor.
TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def relerr(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b), dtype=torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-300))


def assert_close(a, b, tol, what=""):
    e = relerr(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"


def split_levels(flat, counts, trailing):
    out, pos = [], 0
    for c in counts:
        out.append(torch.from_numpy(np.ascontiguousarray(flat[pos:pos + int(c)])).reshape((int(c),) + tuple(trailing)))
        pos += int(c)
    return out
