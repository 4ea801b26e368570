"""psd_safe_cholesky stand-in: torch.linalg.cholesky_ex, then a 3-step diagonal jitter
ladder (1e-6 fp32 / 1e-8 fp64, x10 per try) before raising NotPSDError.  On positive
definite input this is exactly torch.linalg.cholesky_ex, which is all the parity tests use."""
import torch
from .errors import NanError, NotPSDError


def psd_safe_cholesky(A, upper=False, out=None, jitter=None, max_tries=3):
    L, info = torch.linalg.cholesky_ex(A)
    if not torch.any(info):
        return L.mT if upper else L
    if torch.isnan(A).any():
        raise NanError("cholesky of a matrix with NaNs")
    if jitter is None:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    Aj = A.clone()
    prev = 0.0
    for i in range(max_tries):
        new = jitter * (10 ** i)
        Aj.diagonal(dim1=-2, dim2=-1).add_(new - prev)
        prev = new
        L, info = torch.linalg.cholesky_ex(Aj)
        if not torch.any(info):
            return L.mT if upper else L
    raise NotPSDError("Matrix not positive definite after repeatedly adding jitter")
