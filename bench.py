#!/usr/bin/env python
"""Benchmark of the CR hot path: loglik + gradient (mahal_and_det forward + hand-written
backward) block-rows/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): B = 1024 independent LEG series x n = 10^4 block rows,
l = 8, fp32, synthetic posterior-precision blocks; every rank holds its own 1024
series (batch-sharded, weak scaling; the only collective is the all-reduce of the summed
log-likelihood scalar).  A "step" = one forward + backward pass over the rank's batch.
One JSON line is printed by rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version banner there)

import torch  # noqa: E402

WORKLOAD = dict(batch=1024, n=10_000, ell=8, dtype="float32")
METRIC = "CR loglik+grad block-rows/s"
UNIT = "block-rows/s"


def bytes_per_row_level(ell, s):
    """Algorithmic bytes per block-row of ONE level launch: forward reads
    R,O,y and writes D,F,G,R~,O~,x_k,y~; backward reads D,F,G,S~d,S~o,x_k,w~ and writes Sd,So,w.
    Both are (4.5 l^2 + 2 l) * s per input row of that level."""
    return (4.5 * ell * ell + 2 * ell) * s


def bytes_per_row_total(ell, s):
    return (18 * ell * ell + 8 * ell) * s    # fwd + bwd over all levels, per original block-row


def kernel_family(ell, s):
    """Which kernel family libcrb200 dispatches to."""
    if s * ell * ell <= 400:
        return "tpn"                       # thread-per-node (CRB200_TPN_MAX_BLOCK_BYTES)
    return "cs" if (s == 8 and ell == 8) else "level"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, busy=None):
        """`busy`: callable that keeps the GPU under the benchmark's load; nvidia-smi needs a few hundred ms to
        deliver its first line, so short runs repeat the workload until at least two samples exist (<= 3 s)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t0 = time.perf_counter()
        while busy is not None and len(self.lines) < 2 and time.perf_counter() - t0 < 3.0:
            busy()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class LaunchTrace:
    """CUDA-event pair around every native launch (on the launching stream)."""

    def __init__(self):
        self.records = []
        self.enabled = False

    def begin(self, kind, dtype, ell, batch, m):
        if not self.enabled:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        return (kind, m, batch, a, b)

    def end(self, tok):
        if tok is None:
            return
        tok[4].record()
        self.records.append(tok)

    def summary(self):
        agg = {}
        for kind, m, batch, a, b in self.records:
            k = (kind, m, batch)
            t = a.elapsed_time(b)
            cnt, tot = agg.get(k, (0, 0.0))
            agg[k] = (cnt + 1, tot + t)
        return agg


def make_inputs(dev, seed, batch, n, ell, dtype):
    from cyclic_gps.synth import leg_params, leg_precision_blocks
    G, Bm, LLT = leg_params(ell, seed=0, device=dev)
    gen = torch.Generator(device=dev).manual_seed(seed)
    gaps = -torch.log(torch.rand((batch, n - 1), generator=gen, dtype=torch.float64, device=dev)) + 0.01
    R, O = leg_precision_blocks(gaps, G, Bm, LLT, dtype)
    x = torch.randn((batch, n, ell), generator=gen, dtype=dtype, device=dev)
    return R, O, x


def cpu_reference_rows_per_s(n, ell, dtype, series, repeats=1, warm=1):
    """The reference's CPU path (oracle port: same ATen ops + torch autograd backward) on the
    host cores: mahal_and_det fwd+bwd looped over `series` series (the reference has no batch axis)."""
    from oracle import cr_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    G, Bm, LLT = orc.leg_params(ell, seed=0)
    gen = torch.Generator().manual_seed(123)
    items = []
    for _ in range(min(series, 4)):
        gaps = -torch.log(torch.rand(n - 1, generator=gen, dtype=torch.float64)) + 0.01
        R, O = orc.leg_posterior_precision(gaps, G, Bm, LLT)
        x = torch.randn((n, ell), generator=gen, dtype=torch.float64)
        items.append((R.to(dtype), O.to(dtype), x.to(dtype)))
    for i in range(warm):
        orc.loglik_grads_autograd(*items[i % len(items)])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for i in range(series):
            orc.loglik_grads_autograd(*items[i % len(items)])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return series * n / best, best


def workload_name(B, n, ell, dtype_name):
    return f"configs[1]: batched LEG loglik+grad, {B} series x n={n}, l={ell}, {dtype_name}, batch-sharded"


def run_reference(args):
    """--impl reference: the reference CPU implementation of the path (oracle port of
    cyclic_reduction.py + torch autograd, all host threads), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, ell = WORKLOAD["n"], WORKLOAD["ell"]
    dtype = getattr(torch, WORKLOAD["dtype"])
    series_per_step = 8
    cpu_reference_rows_per_s(n, ell, dtype, series_per_step * max(args.warmup, 1), warm=0)
    t_tot = 0.0
    for _ in range(args.steps):
        _, dt = cpu_reference_rows_per_s(n, ell, dtype, series_per_step, warm=0)
        t_tot += dt
    value = args.steps * series_per_step * n / t_tot
    cores = torch.get_num_threads()
    sample = f"{series_per_step} series x n={n} per step (of {WORKLOAD['batch']}), looped one series at a time"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(WORKLOAD["batch"], n, ell, WORKLOAD["dtype"]),
                       "batch_per_gpu": WORKLOAD["batch"], "n": n, "ell": ell,
                       "reference_sample": f"{series_per_step} of {WORKLOAD['batch']} series per step, reference CPU path (oracle port), "
                                           "one series at a time"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_long(args):
    """BASELINE configs[3]: ONE series of n rows (default 1e8), l = 4, fp32, rows spread over the ranks
    (chunk-partitioned CR, cyclic_gps.distributed); strong scaling.  step = loglik forward + backward."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from cyclic_gps import _native, distributed as D
    from cyclic_gps.synth import gaps_for_rows, leg_params, leg_precision_rows
    _native.load()
    n, ell = args.n, args.ell
    dtype = getattr(torch, args.dtype)
    s = torch.empty((), dtype=dtype).element_size()
    plan = D.make_plan(n, world, sub=args.sub)
    lo, hi = plan.rows(rank)
    G, Bm, LLT = leg_params(ell, seed=0, device=dev)
    R, Oprev = leg_precision_rows(gaps_for_rows(lo, hi, n, seed=7, device=dev), G, Bm, LLT, dtype)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn((hi - lo, ell), generator=gen, dtype=dtype, device=dev)
    R.requires_grad_(True); Oprev.requires_grad_(True); x.requires_grad_(True)

    def step():
        R.grad = Oprev.grad = x.grad = None
        mh, ld = D.chunked_mahal_and_det(R, Oprev, x, plan, rank)
        ll = -0.5 * (mh + ld)
        ll.backward()
        return ll

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3) + 2):          # two extra: the caching allocator needs them to settle at this size
        step()
    trace = LaunchTrace()
    _native.TRACE = trace
    clocks = ClockSampler(local)
    sync()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lc0 = _native.launch_count()
    e0.record()
    for _ in range(args.steps):
        ll = step().detach()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    timed_launches = _native.launch_count() - lc0       # kernels of libcrb200 launched inside the timed region
    trace.enabled = True
    for _ in range(min(args.steps, 2)):
        step()
    sync()
    trace.enabled = False
    def _busy():
        step()
        torch.cuda.synchronize()
    clk = clocks.stop(busy=_busy if world == 1 else None) if rank == 0 else None   # (extra steps on one rank only would hang a collective)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if rank != 0:
        dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    agg = trace.summary()
    table = []
    for (kind, m, batch), (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        avg_ms = tot / cnt
        algo = bytes_per_row_level(ell, s) * m * batch
        table.append({"kernel": f"cr_{kernel_family(ell, s)}_{kind}_kernel<{args.dtype},{ell}>", "m": m, "batch": batch, "launches": cnt, "avg_ms": avg_ms,
                      "algo_bytes": algo, "achieved_gbs": algo / (avg_ms * 1e-3) / 1e9, "frac": algo / (avg_ms * 1e-3) / 1e9 / peak})
    top = table[0]
    whole = bytes_per_row_total(ell, s) * (hi - lo)
    roofline = {"bound": "hbm", "kernel": top["kernel"] + f" @ m={top['m']} x batch {top['batch']}", "achieved": top["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": top["frac"], "traffic": None, "peak_source": peak_src,
                "whole_step_rank0": {"algo_bytes": whole, "achieved": whole / (ms / args.steps * 1e-3) / 1e9,
                                     "frac": whole / (ms / args.steps * 1e-3) / 1e9 / peak}}
    if args.levels_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.levels_out)), exist_ok=True)
        with open(args.levels_out, "w") as fh:
            json.dump(table, fh, indent=1)
    cpu = None
    if not args.no_cpu_baseline:
        ncpu = min(n, 1_000_000)
        v, dt = cpu_reference_rows_per_s(ncpu, ell, dtype, 3, warm=1)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"3 passes over a series of n={ncpu} (of {n}; host RAM bounds the reference at ~628 B/row), l={ell}, {args.dtype}, {dt:.1f} s"}
    line = {"metric": METRIC, "value": n * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
            "config": {"workload": f"configs[3]: single long series n={n}, l={ell}, {args.dtype}, chunk-partitioned CR + one all-gather",
                       "n": n, "ell": ell, "sub_chunk_rows": plan.sub, "sub_chunks": plan.nsub, "parallelism": f"row-chunks x{world}",
                       "l2": "inputs per step exceed L2; no explicit flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": None, "gpu_launches": timed_launches,
            "clocks": clk, "loglik": float(ll)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun children) write to fd 1; the driver wants ONE JSON line there.
    From here on fd 1 points at stderr and `emit` writes the line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch"])
    ap.add_argument("--n", type=int, default=WORKLOAD["n"])
    ap.add_argument("--ell", type=int, default=WORKLOAD["ell"])
    ap.add_argument("--dtype", default=WORKLOAD["dtype"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--levels-out", default=None, help="write the per-level launch table (JSON) here")
    ap.add_argument("--workload", default="batch", choices=["batch", "long"],
                    help="batch = configs[1] (default, the headline metric); long = configs[3], one series of --n rows")
    ap.add_argument("--sub", type=int, default=None, help="long workload: rows per sub-chunk (power of two)")
    ap.add_argument("--variant", type=int, default=0, help="force a kernel family (0 auto, 1 lane-per-row, 2 thread-per-node, 3 column-split)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "long":
        if args.n == WORKLOAD["n"]:
            args.n, args.ell = 100_000_000, 4
        return run_long(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from cyclic_gps import _native, cyclic_reduction as cr
    _native.load()
    _native.VARIANT = args.variant
    B, n, ell = args.batch, args.n, args.ell
    dtype = getattr(torch, args.dtype)
    s = torch.empty((), dtype=dtype).element_size()
    R, O, x = make_inputs(dev, 1000 + rank, B, n, ell, dtype)
    Rr, Or, xr = R.requires_grad_(True), O.requires_grad_(True), x.requires_grad_(True)
    total = torch.zeros((), dtype=torch.float64, device=dev)

    def step():
        Rr.grad = Or.grad = xr.grad = None
        mm, dd = cr.mahal_and_det(Rr, Or, xr)
        ll = -0.5 * (mm.double().sum() + dd.double().sum())
        ll.backward()
        tot = ll.detach().clone()
        if dist is not None:
            dist.all_reduce(tot)          # the job-wide log-likelihood (the only collective of this workload)
        return tot

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    trace = LaunchTrace()
    _native.TRACE = trace
    clocks = ClockSampler(local)
    sync()
    if rank == 0:
        clocks.start()
    # pass 1: the headline timing (no per-launch events inside)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lc0 = _native.launch_count()
    e0.record()
    for _ in range(args.steps):
        total += step().detach()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    timed_launches = _native.launch_count() - lc0       # kernels of libcrb200 launched inside the timed region
    # pass 2: same steps with an event pair around every launch (per-kernel durations, roofline)
    trace.enabled = True
    for _ in range(args.steps):
        step()
    sync()
    trace.enabled = False
    def _busy():
        step()
        torch.cuda.synchronize()
    clk = clocks.stop(busy=_busy if world == 1 else None) if rank == 0 else None   # (extra steps on one rank only would hang a collective)
    launches_per_step = len(trace.records) // max(args.steps, 1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    rows = B * n * world
    value = rows * args.steps / (ms * 1e-3)

    # ---- end to end: host buffers in, scalars out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hR = torch.empty(R.shape, dtype=dtype, pin_memory=True).copy_(R.detach())
        hO = torch.empty(O.shape, dtype=dtype, pin_memory=True).copy_(O.detach())
        hx = torch.empty(x.shape, dtype=dtype, pin_memory=True).copy_(x.detach())
        hout = torch.empty((2, B), dtype=dtype, pin_memory=True)
        dR, dO, dx = torch.empty_like(R), torch.empty_like(O), torch.empty_like(x)

        copy_stream = torch.cuda.Stream(device=dev)
        nchunk = 8 if B % 8 == 0 and B >= 64 else 1
        bsz = B // nchunk

        def e2e_step():
            """Host buffers in, per-series scalars out, through the public API (cr.mahal_and_det + backward).
            The batch is cut into chunks of series; chunk c+1 crosses PCIe on a copy stream while chunk c computes."""
            main = torch.cuda.current_stream()
            copy_stream.wait_stream(main)                 # previous step is done with the device buffers
            ready = []
            with torch.cuda.stream(copy_stream):
                for c in range(nchunk):
                    sl = slice(c * bsz, (c + 1) * bsz)
                    dR[sl].copy_(hR[sl], non_blocking=True)
                    dO[sl].copy_(hO[sl], non_blocking=True)
                    dx[sl].copy_(hx[sl], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    ready.append(ev)
            tot = torch.zeros((), dtype=torch.float64, device=dev)
            for c in range(nchunk):
                sl = slice(c * bsz, (c + 1) * bsz)
                main.wait_event(ready[c])
                Rc, Oc, xc = (t[sl].detach().requires_grad_(True) for t in (dR, dO, dx))
                mm, dd = cr.mahal_and_det(Rc, Oc, xc)
                ll = -0.5 * (mm.double().sum() + dd.double().sum())
                ll.backward()
                tot += ll.detach()
                hout[0, sl].copy_(mm.detach(), non_blocking=True)
                hout[1, sl].copy_(dd.detach(), non_blocking=True)
            if dist is not None:
                dist.all_reduce(tot)
            main.synchronize()                            # the caller holds the result on the host
            return hout

        e2e_step()
        sync()
        k2 = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(k2):
            e2e_step()
        a1.record()
        sync()
        ms2 = a0.elapsed_time(a1) / k2
        if dist is not None:
            t = torch.tensor([ms2], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t)
        h2d = (hR.numel() + hO.numel() + hx.numel()) * s
        e2e = {"value": rows / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": hout.numel() * s,
               "ms_per_step": ms2, "steps": k2, "pipeline": f"{nchunk} chunks of {bsz} series, H2D on a copy stream overlapped with compute",
               "h2d_gbs": h2d / (ms2 * 1e-3) / 1e9}
        del hR, hO, hx, dR, dO, dx

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch CUDA-event durations)
    peak, peak_src = peaks()
    agg = trace.summary()
    table = []
    fam = kernel_family(ell, s)
    for (kind, m, batch), (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        avg_ms = tot / cnt
        algo = bytes_per_row_level(ell, s) * m * batch
        table.append({"kernel": f"cr_{fam}_{kind}_kernel<{args.dtype},{ell}>", "kind": kind, "m": m, "batch": batch, "launches": cnt,
                      "avg_ms": avg_ms, "algo_bytes": algo, "achieved_gbs": algo / (avg_ms * 1e-3) / 1e9,
                      "frac": algo / (avg_ms * 1e-3) / 1e9 / peak})
    top = table[0]
    step_kernel_ms = sum(r["avg_ms"] * r["launches"] for r in table) / max(args.steps, 1)
    roofline = {"bound": "hbm", "kernel": top["kernel"] + f" @ m={top['m']}", "achieved": top["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": top["frac"], "traffic": None, "peak_source": peak_src,
                "algo_bytes_per_launch": top["algo_bytes"], "avg_launch_ms": top["avg_ms"],
                "share_of_step": top["avg_ms"] / step_kernel_ms if step_kernel_ms else None,
                "whole_step": {"algo_bytes": bytes_per_row_total(ell, s) * B * n,
                               "achieved": bytes_per_row_total(ell, s) * B * n / (ms / args.steps * 1e-3) / 1e9,
                               "frac": bytes_per_row_total(ell, s) * B * n / (ms / args.steps * 1e-3) / 1e9 / peak}}
    # DRAM bytes of that kernel from the committed ncu --set full capture, scaled to this launch's rows
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                ent = json.load(fh).get(top["kernel"].split("<")[0] + f"@{args.dtype},{ell}")
            if ent:
                roofline["traffic"] = ent["bytes_per_row"] * top["m"] * top["batch"]
                roofline["traffic_source"] = "profiles/traffic.json (ncu dram__bytes_read+write per block-row of the level-0 launch, x rows of this launch)"
        except Exception:
            pass
    if args.levels_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.levels_out)), exist_ok=True)
        with open(args.levels_out, "w") as fh:
            json.dump(table, fh, indent=1)

    cpu = None
    if not args.no_cpu_baseline:
        _, t_probe = cpu_reference_rows_per_s(n, ell, dtype, 4)            # calibrate: aim at ~12 s of CPU work
        series = int(max(8, min(B, 12.0 / max(t_probe / 4, 1e-4))))
        v, dt = cpu_reference_rows_per_s(n, ell, dtype, series, warm=0)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{series} of {B} series (n={n}, l={ell}, {args.dtype}), one series at a time (the reference has no batch axis), "
                         f"oracle port of the reference + torch autograd, {dt:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
            "config": {"workload": workload_name(B, n, ell, args.dtype),
                       "batch_per_gpu": B, "n": n, "ell": ell, "parallelism": f"batch-shard x{world}",
                       "l2": "inputs per step (%.1f GB) exceed L2 (126 MB); no explicit flush" % ((R.numel() + O.numel() + x.numel()) * s / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": timed_launches,
            "gpu_launches_per_step": timed_launches // max(args.steps, 1), "per_level_table_launches_per_step": launches_per_step, "clocks": clk,
            "loglik_checksum": float(total) / max(args.steps, 1)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
