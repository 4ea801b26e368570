"""Stand-in for pytorch_lightning (oracle scaffolding): just enough for LEGFamily to import."""
import torch


class LightningModule(torch.nn.Module):
    def log(self, *a, **k):
        return None


class Trainer:
    def __init__(self, *a, **k):
        raise RuntimeError("pytorch_lightning is not installed; Trainer is a stub")
