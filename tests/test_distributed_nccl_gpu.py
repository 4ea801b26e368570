"""GPU (>= 2 devices): the chunk-partitioned long-series path with the boundary system exchanged over NCCL, one rank
per GPU under torchrun (tests/nccl_worker.py).  Skipped on a single-GPU box; run through `gpurun --gpus 2` (log under
profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_chunked_two_ranks_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL ranks must not share a device)")
    port = 29900 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert r.stdout.count("NCCL_PARITY_OK") == 2, r.stdout[-4000:]
