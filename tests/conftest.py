"""pytest configuration: registers the `gpu` marker and puts the product package
(`cyclic-gps_b200/`, which holds the importable `cyclic_gps` package) and the repo root
(for `oracle/`) on sys.path.  `/root/reference` is never used at test time."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cyclic-gps_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")
    return {name: np.load(os.path.join(d, name + ".npz"), allow_pickle=False)
            for name in ("random_llt", "known", "leg", "helpers", "leg_model", "predictions")}
