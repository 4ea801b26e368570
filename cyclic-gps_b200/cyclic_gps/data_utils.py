"""Synthetic series and the single-series dataset wrapper, with the reference's names
(cyclic_gps/data_utils.py:44-75).  Input generation only."""
import scipy.ndimage
import torch
from torch.utils.data import Dataset


def generate_data(num_datapoints, data_dim, data_type, spacing: str = "irregular"):
    """(ts:(n,), xs:(n,d)): exponential(1)+0.01 gaps or unit gaps; values are white noise
    smoothed by a Gaussian filter of width 10 (reference data_utils.py:44-57)."""
    if spacing == "irregular":
        gaps = torch.empty(num_datapoints, dtype=data_type).exponential_(1.0) + 0.01
        ts = torch.cumsum(gaps, dim=0)
    else:
        ts = torch.cumsum(torch.ones(num_datapoints), dim=0)
    cols = []
    for _ in range(data_dim):
        noise = torch.randn(num_datapoints, dtype=data_type).numpy()
        cols.append(torch.from_numpy(scipy.ndimage.gaussian_filter1d(noise, 10, axis=0)).reshape(-1, 1))
    return ts, torch.cat(cols, dim=-1)


class time_series_dataset(Dataset):
    """Batch of one: always serves series 0 (reference data_utils.py:61-75)."""

    def __init__(self, ts, xs):
        self.ts, self.xs = ts, xs

    def __len__(self):
        return self.ts.shape[0]

    def __getitem__(self, idx):
        return self.ts[0, :], self.xs[0, :, :]
