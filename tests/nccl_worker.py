"""Worker of tests/test_distributed_nccl_gpu.py, launched by torchrun with one rank per GPU: the chunk-partitioned
long-series path with the boundary system exchanged over NCCL (all_gather_into_tensor on real devices), checked on
every rank against the CPU oracle of the unchunked algorithm and, at a larger size, against the plain single-sweep
GPU path.  Prints one line `NCCL_PARITY_OK ...` per rank on success."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert dist.get_backend() == "nccl"
    from test_distributed_gpu import _run
    import bench
    # small sizes against the CPU oracle (tolerances of helpers.TOL), several shapes incl. a ragged tail and an empty rank
    cases = [(6000, 4, torch.float32, 256), (2500, 3, torch.float64, 128), (4097, 8, torch.float32, 512),
             (777, 16, torch.float64, 32), (100, 2, torch.float64, 64)]
    for (n, l, dtype, sub) in cases:
        _run(rank, world, n, l, dtype, sub)
    # a larger series against the single-sweep GPU path (bench.py's parity block) + the residual of J w = x
    rk = bench.Ranks.__new__(bench.Ranks)
    rk.rank, rk.world, rk.local, rk.dev, rk.dist = rank, world, local, torch.device("cuda", local), dist
    par = bench.long_parity(rk, 4, torch.float32, 400_000, None)
    if rank == 0:
        assert par["ok"], par
    dist.barrier()
    print(f"NCCL_PARITY_OK rank={rank} world={world} cases={len(cases)} parity={par}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
