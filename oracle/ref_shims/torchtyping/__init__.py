"""Stand-in for torchtyping: annotations only, no runtime checks (oracle scaffolding)."""
import torch


class _TT:
    def __getitem__(self, item):
        return torch.Tensor

    def __call__(self, *a, **k):
        return torch.Tensor


TensorType = _TT()


def patch_typeguard():
    return None
