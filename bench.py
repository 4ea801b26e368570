#!/usr/bin/env python
"""Benchmark of the CR hot path: loglik + gradient (mahal_and_det forward + hand-written
backward) block-rows/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): B = 1024 independent LEG series x n = 10^4 block rows,
l = 8, fp32, synthetic posterior-precision blocks; every rank holds its own 1024
series (batch-sharded, weak scaling; the only collective is the all-reduce of the summed
log-likelihood scalar).  A "step" = one forward + backward pass over the rank's batch.
One JSON line is printed by rank 0.  Besides the headline (`value`, weak scaling) the line carries
  strong       the same workload with 1024 series in TOTAL (1024 / N per GPU),
  long_series  BASELINE configs[3]: ONE series of n = 1e8 rows, l = 4, fp32, chunk-partitioned over the N ranks with one
               NCCL all-gather (cyclic_gps.distributed), with its own clocks record and a parity block (chunked on N
               ranks vs the unchunked single sweep at n = 4e6, and the residual of J w = x at the full size)."""
import argparse
import gc
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
    if s == 8 and ell == 8:
        return "cs"                        # column-split
    return "mma" if (ell >= 10 if s == 8 else ell >= 17) else "level"    # warp-per-node DMMA / lane-per-row (cr_inst.cu kMmaAuto)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def quiet_gc():
    """Collect now and keep the cyclic collector off during a timed loop (as timeit does); returns the previous state."""
    was = gc.isenabled()
    gc.collect()
    gc.disable()
    return was


def restore_gc(was):
    if was:
        gc.enable()


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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
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


def keep_busy_until_sampled(sampler, step, rank, world, dist, dev, min_samples=5, max_s=8.0):
    """nvidia-smi needs a few hundred ms per line; short timed regions are followed by the same steps, run by ALL
    ranks together (rank 0 decides when to stop and broadcasts it), until rank 0 holds `min_samples` lines taken
    under the benchmark's own load."""
    t0 = time.perf_counter()
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    while True:
        step()
        torch.cuda.synchronize()
        done = sampler is None or sampler.proc is None or len(sampler.lines) >= min_samples or time.perf_counter() - t0 > max_s
        if world > 1:
            flag.fill_(1 if done else 0)
            dist.broadcast(flag, 0)
            done = bool(int(flag.item()))
        if done:
            return


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


def batch_config(B, n, ell, dtype_name, world):
    """`config` of the JSON line; the reference arm prints the very same dict (the driver compares them)."""
    s = 4 if dtype_name == "float32" else 8
    in_bytes = (B * n * ell * ell + B * (n - 1) * ell * ell + B * n * ell) * s
    return {"workload": workload_name(B, n, ell, dtype_name), "batch_per_gpu": B, "n": n, "ell": ell,
            "parallelism": f"batch-shard x{world}",
            "l2": "inputs per step (%.1f GB) exceed L2 (126 MB); no explicit flush" % (in_bytes / 1e9)}


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
            "config": batch_config(WORKLOAD["batch"], n, ell, WORKLOAD["dtype"], args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample + "; reference CPU path = oracle port of cyclic_reduction.py + torch autograd"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


class Ranks:
    """Process-group plumbing shared by the workloads of one bench.py run."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def sync(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def long_rows(lo, hi, n, ell, dtype, dev, seed):
    """Rows [lo, hi) of THE global long series (gaps and right-hand side are functions of the global row index
    only, so any partition over ranks sees the same series)."""
    from cyclic_gps.synth import gaps_for_rows, leg_params, leg_precision_rows, x_for_rows
    G, Bm, LLT = leg_params(ell, seed=0, device=dev)
    R, Oprev = leg_precision_rows(gaps_for_rows(lo, hi, n, seed=seed, device=dev), G, Bm, LLT, dtype)
    return R, Oprev, x_for_rows(lo, hi, ell, seed + 1, dev, dtype)


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300)) if b.numel() else 0.0


def long_parity(rk, ell, dtype, n_par, sub):
    """Chunk-partitioned path on `world` ranks (NCCL all-gather when world > 1) against the plain single-sweep path
    of the same library on rank 0, same global series: mahal, logdet and the gradients of the rows rank 0 owns."""
    from cyclic_gps import cyclic_reduction as cr, distributed as D
    plan = D.make_plan(n_par, rk.world, sub=sub)
    lo, hi = plan.rows(rk.rank)
    R, Oprev, x = long_rows(lo, hi, n_par, ell, dtype, rk.dev, seed=11)
    R.requires_grad_(True); Oprev.requires_grad_(True); x.requires_grad_(True)
    mh, ld = D.chunked_mahal_and_det(R, Oprev, x, plan, rk.rank)
    (mh + ld).backward()
    out = None
    if rk.rank == 0:
        Rf, Of, xf = long_rows(0, n_par, n_par, ell, dtype, rk.dev, seed=11)
        Rf.requires_grad_(True); xf.requires_grad_(True)
        Os = Of[1:].clone().requires_grad_(True)
        m1, d1 = cr.mahal_and_det(Rf, Os, xf)
        (m1 + d1).backward()
        gOprev = torch.cat([torch.zeros_like(Os.grad[:1]), Os.grad], dim=0)
        out = {"n": n_par, "ranks": rk.world, "backend": "nccl" if rk.world > 1 else "single process",
               "sub_chunk_rows": plan.sub, "rows_checked_for_gradients": hi - lo,
               "mahal_rel": _rel(mh, m1), "logdet_rel": _rel(ld, d1),
               "gR_rel": _rel(R.grad, Rf.grad[lo:hi]), "gO_rel": _rel(Oprev.grad, gOprev[lo:hi]), "gx_rel": _rel(x.grad, xf.grad[lo:hi]),
               "tolerance": 1e-4}
        out["ok"] = all(out[k] <= out["tolerance"] for k in ("mahal_rel", "logdet_rel", "gR_rel", "gO_rel", "gx_rel"))
    rk.sync()
    return out


@torch.no_grad()
def long_residual(rk, R, Oprev, x, w, chunk=1 << 22):
    """max |J w - x| / max |x| over ALL rows of the distributed series: every rank needs w of the row before its first
    and after its last row, and the coupling block of the row after its last (one tiny all-gather)."""
    ell = R.shape[-1]
    n_loc = R.shape[0]
    z = lambda *sh: torch.zeros(sh, dtype=torch.float64, device=rk.dev)
    mine = torch.cat([(w[:1].double().reshape(-1) if n_loc else z(ell)), (w[-1:].double().reshape(-1) if n_loc else z(ell)),
                      (Oprev[:1].double().reshape(-1) if n_loc else z(ell * ell)), torch.tensor([float(n_loc > 0)], dtype=torch.float64, device=rk.dev)])
    if rk.dist is not None:
        allv = torch.empty((rk.world, mine.numel()), dtype=torch.float64, device=rk.dev)
        rk.dist.all_gather_into_tensor(allv, mine.unsqueeze(0))
    else:
        allv = mine.unsqueeze(0)
    prev = [r for r in range(rk.rank) if allv[r, -1] > 0]
    nxt = [r for r in range(rk.rank + 1, rk.world) if allv[r, -1] > 0]
    w_before = allv[prev[-1], ell:2 * ell] if prev else z(ell)
    w_after = allv[nxt[0], 0:ell] if nxt else z(ell)
    O_after = allv[nxt[0], 2 * ell:2 * ell + ell * ell].view(ell, ell) if nxt else z(ell, ell)
    worst = z(1)
    xmax = x.detach().abs().max().double().reshape(1) if n_loc else z(1)
    for a in range(0, n_loc, chunk):
        b = min(a + chunk, n_loc)
        wd = w[a:b].double()
        r = torch.einsum("nij,nj->ni", R[a:b].detach().double(), wd) - x[a:b].detach().double()
        wl = torch.cat([(w[a - 1:a].double() if a > 0 else w_before.view(1, ell)), wd[:-1]], dim=0)
        r += torch.einsum("nij,nj->ni", Oprev[a:b].detach().double(), wl)
        On = torch.cat([Oprev[a + 1:b].detach().double(), (Oprev[b:b + 1].detach().double() if b < n_loc else O_after.view(1, ell, ell))], dim=0)
        wr = torch.cat([wd[1:], (w[b:b + 1].double() if b < n_loc else w_after.view(1, ell))], dim=0)
        r += torch.einsum("nji,nj->ni", On, wr)
        worst = torch.maximum(worst, r.abs().max().reshape(1))
    if rk.dist is not None:
        rk.dist.all_reduce(worst, op=rk.dist.ReduceOp.MAX)
        rk.dist.all_reduce(xmax, op=rk.dist.ReduceOp.MAX)
    return float(worst / xmax)


def run_long(args, rk, standalone):
    """BASELINE configs[3]: ONE series of n rows (default 1e8), l = 4, fp32, rows spread over the ranks
    (chunk-partitioned CR + one NCCL all-gather, cyclic_gps.distributed); strong scaling.  step = loglik forward +
    backward.  Returns the result dict on rank 0 (None elsewhere)."""
    from cyclic_gps import _native, distributed as D
    rank, world, dev, dist = rk.rank, rk.world, rk.dev, rk.dist
    n, ell = args.long_n, args.long_ell
    dtype = getattr(torch, args.long_dtype)
    s = torch.empty((), dtype=dtype).element_size()
    D.DEFERRED_PD_CHECK = True                 # this loop always runs backward(): no GPU drain between the passes
    D.RELEASE_FACTORS_AFTER_BACKWARD = True
    parity = None
    if not args.no_parity:
        parity = long_parity(rk, ell, dtype, min(args.parity_n, n), None)
        torch.cuda.empty_cache()
    plan = D.make_plan(n, world, sub=args.sub)
    lo, hi = plan.rows(rank)
    R, Oprev, x = long_rows(lo, hi, n, ell, dtype, dev, seed=7)
    R.requires_grad_(True); Oprev.requires_grad_(True); x.requires_grad_(True)

    def step():
        R.grad = Oprev.grad = x.grad = None
        mh, ld = D.chunked_mahal_and_det(R, Oprev, x, plan, rank)
        ll = -0.5 * (mh + ld)
        ll.backward()
        return ll

    for _ in range(max(args.warmup, 3) + 2):          # two extra: the caching allocator needs them to settle at this size
        step()
    trace = LaunchTrace()
    _native.TRACE = trace
    clocks = ClockSampler(rk.local) if rank == 0 else None
    rk.sync()
    if clocks is not None:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lc0 = _native.launch_count()
    gc_was = quiet_gc()               # a collector pause on ONE rank stalls all of them at the all-gather (the step is ~3 ms of kernels at 8 GPUs)
    e0.record()
    for _ in range(args.steps):
        ll = step().detach()
    e1.record()
    rk.sync()
    restore_gc(gc_was)
    ms = e0.elapsed_time(e1)
    timed_launches = _native.launch_count() - lc0       # kernels of libcrb200 launched inside the timed region
    trace.enabled = True
    for _ in range(min(args.steps, 2)):
        step()
    rk.sync()
    trace.enabled = False
    keep_busy_until_sampled(clocks, step, rank, world, dist, dev)       # all ranks keep stepping until rank 0 has its samples
    clk = clocks.stop() if clocks is not None else None
    _native.TRACE = None
    ms = rk.max_over_ranks(ms)
    # J w = x at the full size: x.grad of the last step is d(-0.5 mahal)/dx = -w
    residual = None
    if not args.no_parity:
        residual = long_residual(rk, R, Oprev, x, -x.grad)
    if rank != 0:
        return None
    peak, peak_src = peaks()
    agg = trace.summary()
    table = []
    for (kind, m, batch), (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        avg_ms = tot / cnt
        algo = bytes_per_row_level(ell, s) * m * batch
        table.append({"kernel": f"cr_{kernel_family(ell, s)}_{kind}_kernel<{args.long_dtype},{ell}>", "m": m, "batch": batch, "launches": cnt, "avg_ms": avg_ms,
                      "algo_bytes": algo, "achieved_gbs": algo / (avg_ms * 1e-3) / 1e9, "frac": algo / (avg_ms * 1e-3) / 1e9 / peak})
    top = table[0]
    whole = bytes_per_row_total(ell, s) * (hi - lo)
    roofline = {"bound": "hbm", "kernel": top["kernel"] + f" @ m={top['m']} x batch {top['batch']}", "achieved": top["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": top["frac"], "traffic": None, "peak_source": peak_src,
                "whole_step_rank0": {"rows": hi - lo, "algo_bytes": whole, "achieved": whole / (ms / args.steps * 1e-3) / 1e9,
                                     "frac": whole / (ms / args.steps * 1e-3) / 1e9 / peak}}
    if args.levels_out and standalone:
        os.makedirs(os.path.dirname(os.path.abspath(args.levels_out)), exist_ok=True)
        with open(args.levels_out, "w") as fh:
            json.dump(table, fh, indent=1)
    if parity is not None:
        parity["residual_Jw_minus_x_rel_full_n"] = residual
        parity["ok"] = bool(parity["ok"] and residual is not None and residual <= 1e-4)
    return {"workload": f"configs[3]: single long series n={n}, l={ell}, {args.long_dtype}, chunk-partitioned CR + one all-gather",
            "metric": METRIC, "value": n * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "ms_per_step": ms / args.steps, "scaling": "strong", "n": n, "ell": ell, "dtype": "f32" if dtype == torch.float32 else "f64",
            "sub_chunk_rows": plan.sub, "sub_chunks": plan.nsub, "parallelism": f"row-chunks x{world}",
            "collective": "one all_gather_into_tensor (nccl) of the boundary system per forward pass" if world > 1 else "none (one rank)",
            "roofline": roofline, "gpu_launches": timed_launches, "clocks": clk, "parity": parity, "loglik": float(ll)}


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


def bench_model(ell, dtype, dev, train):
    """A LEGFamily whose matrices equal synth.leg_params(ell, seed=0) -- the model behind the synthetic blocks."""
    from cyclic_gps.models import LEGFamily
    from cyclic_gps.synth import leg_params
    G, Bm, LLT = leg_params(ell, seed=0)
    model = LEGFamily(rank=ell, obs_dim=1, train=train, data_type=dtype)
    Rm = torch.tril(0.5 * (G - G.T), diagonal=-1)                     # G = I + Rm - Rm^T + 1e-5 I
    with torch.no_grad():
        model.R_params.copy_(Rm[model.R_idxs].to(dtype))
    model.register_model_matrices_from_params()
    return model


def run_e2e(args, rk, cr, B, n, ell, dtype, R_dev, O_dev):
    """e2e: the headline metric (mahal_and_det forward + backward to gR, gO, gx) fed from host-resident time stamps and
    observations through the public API; train: the whole LEGFamily.log_likelihood step (two factorisations, gradient
    wrt the model parameters), also from host buffers."""
    dev, dist = rk.dev, rk.dist
    s = torch.empty((), dtype=dtype).element_size()
    gen = torch.Generator(device=dev).manual_seed(1000 + rk.rank)      # the same gaps as make_inputs
    gaps = -torch.log(torch.rand((B, n - 1), generator=gen, dtype=torch.float64, device=dev)) + 0.01
    ts = torch.cat([torch.zeros((B, 1), dtype=torch.float64, device=dev), torch.cumsum(gaps, dim=1)], dim=1)
    xs = torch.randn((B, n, 1), generator=gen, dtype=dtype, device=dev)
    h_ts = torch.empty(ts.shape, dtype=torch.float64, pin_memory=True).copy_(ts)
    h_xs = torch.empty(xs.shape, dtype=dtype, pin_memory=True).copy_(xs)
    hout = torch.empty((2, B), dtype=dtype, pin_memory=True)
    d_ts, d_xs = torch.empty_like(ts), torch.empty_like(xs)
    del gaps, ts, xs
    model = bench_model(ell, dtype, dev, train=False)
    _, shift = model._obs_terms()
    rows = B * n * rk.world

    # builder check: the device-built blocks of series 0 against the fp64 torch construction used for the timed inputs
    with torch.no_grad():
        Rb, Ob = model._precision_blocks(h_ts[:1].to(dev), shift)
        chk = max(float((Rb[0] - R_dev[0]).abs().max() / R_dev[0].abs().max()), float((Ob[0] - O_dev[0]).abs().max() / O_dev[0].abs().max()))
        del Rb, Ob

    copy_stream = torch.cuda.Stream(device=dev)
    # chunk plan: a SMALL first chunk (its H2D copy is the only one nothing can hide), the rest in equal parts
    if B >= 64 and args.e2e_chunks > 1:
        first = max(1, B // args.e2e_first_frac)
        rest = args.e2e_chunks - 1
        sizes = [first] + [(B - first) // rest + (1 if i < (B - first) % rest else 0) for i in range(rest)]
    else:
        sizes = [B]
    bounds = [0]
    for z in sizes:
        bounds.append(bounds[-1] + z)
    nchunk = len(sizes)

    def e2e_step():
        """The batch is cut into chunks of series: chunk c + 1 crosses PCIe on a copy stream while chunk c is built and reduced."""
        main_s = torch.cuda.current_stream()
        copy_stream.wait_stream(main_s)                   # the previous step is done with the device buffers
        ready = []
        with torch.cuda.stream(copy_stream):
            for c in range(nchunk):
                sl = slice(bounds[c], bounds[c + 1])
                d_ts[sl].copy_(h_ts[sl], non_blocking=True)
                d_xs[sl].copy_(h_xs[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append(ev)
        tot = torch.zeros((), dtype=torch.float64, device=dev)
        for c in range(nchunk):
            sl = slice(bounds[c], bounds[c + 1])
            main_s.wait_event(ready[c])
            with torch.no_grad():
                Rs, Os = model._precision_blocks(d_ts[sl], shift)
                v = model.compute_v(d_xs[sl])
            Rs.requires_grad_(True); Os.requires_grad_(True); v.requires_grad_(True)
            mm, dd = cr.mahal_and_det(Rs, Os, v)
            ll = -0.5 * (mm.double().sum() + dd.double().sum())
            ll.backward()
            tot += ll.detach()
            hout[0, sl].copy_(mm.detach(), non_blocking=True)
            hout[1, sl].copy_(dd.detach(), non_blocking=True)
        if dist is not None:
            dist.all_reduce(tot)
        main_s.synchronize()                              # the caller holds the result on the host
        return hout

    # second pipeline: the batch is BUILT slice by slice (the builder of slice c runs while slice c + 1 crosses PCIe: builder and copy take
    # about the same time) into one (B, n, l, l) pair of tensors, and ONE cyclic reduction runs over all series (a 128-series sweep costs
    # 18 us per series, the full batch 8.6).  The blocks live in storage the caller owns (peg_precision(out=...)).
    if B >= 64 and args.e2e_chunks > 1:
        first2 = max(1, B // args.e2e_first_frac)
        sizes2 = [first2] + [(B - first2) // 3 + (1 if i < (B - first2) % 3 else 0) for i in range(3)]
    else:
        sizes2 = [B]
    bounds2 = [0]
    for z in sizes2:
        bounds2.append(bounds2[-1] + z)
    R_all = torch.empty((B, n, ell, ell), dtype=dtype, device=dev)
    O_all = torch.empty((B, n - 1, ell, ell), dtype=dtype, device=dev)

    def e2e_step_build_stream():
        main_s = torch.cuda.current_stream()
        copy_stream.wait_stream(main_s)
        ready = []
        with torch.cuda.stream(copy_stream):
            for c in range(len(sizes2)):
                sl = slice(bounds2[c], bounds2[c + 1])
                d_ts[sl].copy_(h_ts[sl], non_blocking=True)
                d_xs[sl].copy_(h_xs[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append(ev)
        with torch.no_grad():
            for c in range(len(sizes2)):
                sl = slice(bounds2[c], bounds2[c + 1])
                main_s.wait_event(ready[c])
                model._precision_blocks(d_ts[sl], shift, out=(R_all[sl], O_all[sl]))
            v = model.compute_v(d_xs)
        Rs, Os = R_all.detach().requires_grad_(True), O_all.detach().requires_grad_(True)
        v.requires_grad_(True)
        mm, dd = cr.mahal_and_det(Rs, Os, v)
        ll = -0.5 * (mm.double().sum() + dd.double().sum())
        ll.backward()
        tot = ll.detach()
        hout[0].copy_(mm.detach(), non_blocking=True)
        hout[1].copy_(dd.detach(), non_blocking=True)
        if dist is not None:
            dist.all_reduce(tot)
        main_s.synchronize()
        return hout

    def timed(fn, k):
        fn(); fn()
        rk.sync()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc_was = quiet_gc()
        a0.record()
        for _ in range(k):
            fn()
        a1.record()
        rk.sync()
        restore_gc(gc_was)
        return rk.max_over_ranks(a0.elapsed_time(a1) / k)

    k2 = max(3, min(args.steps, 5))
    ms2 = timed(e2e_step, k2)
    ref_out = e2e_step().clone()
    pipe = (f"chunks of {sizes} series, H2D on a copy stream under the compute of the previous chunk; "
            "pinned host time stamps (fp64) + observations -> device; precision blocks built on the device "
            "(crb200_peg_precision_fwd = the reference's compute_posterior_precision); cr.mahal_and_det + backward to gR, gO, gx; "
            "per-series scalars -> host")
    other = {"pipeline": "one cyclic reduction over the whole batch, blocks built slice by slice", "ms_per_step": None}
    try:
        ms2b = timed(e2e_step_build_stream, k2)
        same = bool(torch.allclose(e2e_step_build_stream(), ref_out, rtol=1e-5, atol=1e-3))
        other["ms_per_step"], other["same_scalars_as_the_chunked_pipeline"] = ms2b, same
        if same and ms2b < ms2:
            other = {"pipeline": "chunks of series, builder + cyclic reduction per chunk", "ms_per_step": ms2}
            ms2 = ms2b
            pipe = (f"slices of {sizes2} series cross PCIe on a copy stream while the precision builder (crb200_peg_precision_fwd = the reference's "
                    "compute_posterior_precision, peg_precision(out=...)) fills the previous slice's blocks of ONE (B, n, l, l) batch; pinned host time "
                    "stamps (fp64) + observations -> device; then ONE cr.mahal_and_det over all series + backward to gR, gO, gx; per-series scalars -> host")
    except Exception as ex:  # noqa: BLE001
        other["error"] = f"{type(ex).__name__}: {ex}"[:200]
    del R_all, O_all
    h2d = h_ts.numel() * 8 + h_xs.numel() * s
    e2e = {"value": rows / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": hout.numel() * s,
           "ms_per_step": ms2, "steps": k2, "pipeline": pipe, "other_pipeline": other,
           "builder_vs_fp64_torch_blocks_rel": chk}

    # the whole training step of the reference (LEGFamily.log_likelihood: prior log-det + posterior mahal / log-det,
    # gradient wrt N, R, B, Lambda) from the same host buffers
    tmodel = bench_model(ell, dtype, dev, train=True)
    hgrad = torch.empty(tmodel.parameter_count, dtype=dtype, pin_memory=True)

    def train_step():
        d_ts.copy_(h_ts, non_blocking=True)
        d_xs.copy_(h_xs, non_blocking=True)
        for p in tmodel.parameters():
            p.grad = None
        ll = tmodel.log_likelihood(d_ts, d_xs).sum()
        ll.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in tmodel.parameters()])
        if dist is not None:
            flat = flat.to(dev)
            dist.all_reduce(flat)
        hgrad.copy_(flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return ll

    try:
        ms3 = timed(train_step, k2)
        train = {"what": "LEGFamily.log_likelihood(ts, xs).sum().backward(): device precision builder (posterior blocks + the prior's "
                         "log-determinant in one kernel, hand-written backward) + ONE CR factorisation (the reference runs two), "
                         "gradient of the log-likelihood wrt the model parameters", "ms_per_step": ms3,
                 "value": rows / (ms3 * 1e-3), "unit": UNIT, "parameters": int(tmodel.parameter_count),
                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": hgrad.numel() * s}
    except Exception as ex:  # noqa: BLE001
        train = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    return e2e, train


def time_graphed(rk, tensors, steps, warmup):
    """The same loglik+grad step through cyclic_gps.graphs.GraphedMahalAndDet: ONE CUDA-graph replay per step instead of Python,
    autograd and allocator work around ~12-18 launches (what bounds the strong-scaling step at 128 series per GPU).  Returns ms
    for `steps` steps, max over ranks, or None if the graph could not be built."""
    from cyclic_gps.graphs import GraphedMahalAndDet
    dist = rk.dist
    try:
        gr = GraphedMahalAndDet(*[t.detach() for t in tensors], g_mahal=-0.5, g_det=-0.5)
    except Exception:  # noqa: BLE001
        return None

    def step():
        mh, ld, gR, gO, gx = gr(gr.R, gr.O, gr.x)              # static buffers: nothing is copied, the gradients land in gR / gO / gx
        tot = -0.5 * (mh.double().sum() + ld.double().sum())
        if dist is not None:
            dist.all_reduce(tot)
        return tot

    for _ in range(warmup):
        step()
    rk.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc_was = quiet_gc()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    rk.sync()
    restore_gc(gc_was)
    gr.check()
    ms = rk.max_over_ranks(e0.elapsed_time(e1))
    del gr
    return ms


def time_batch(rk, cr, tensors, steps, warmup):
    """K timed loglik+grad steps over this rank's series; returns (ms max over ranks, step fn, checksum, launches)."""
    from cyclic_gps import _native
    Rr, Or, xr = tensors
    dist = rk.dist
    total = torch.zeros((), dtype=torch.float64, device=rk.dev)

    def step():
        Rr.grad = Or.grad = xr.grad = None
        mm, dd = cr.mahal_and_det(Rr, Or, xr)
        ll = -0.5 * (mm.double().sum() + dd.double().sum())
        ll.backward()
        tot = ll.detach().clone()
        if dist is not None:
            dist.all_reduce(tot)          # the job-wide log-likelihood (the only collective of this workload)
        return tot

    for _ in range(warmup):
        step()
    rk.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lc0 = _native.launch_count()
    gc_was = quiet_gc()
    e0.record()
    for _ in range(steps):
        total += step().detach()
    e1.record()
    rk.sync()
    restore_gc(gc_was)
    ms = rk.max_over_ranks(e0.elapsed_time(e1))
    return ms, step, float(total) / max(steps, 1), _native.launch_count() - lc0


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
    ap.add_argument("--e2e-chunks", type=int, default=2, help="e2e: chunks of series whose H2D copies overlap the previous chunk's compute")
    ap.add_argument("--e2e-first-frac", type=int, default=8, help="e2e: the first chunk holds batch / this many series (its copy cannot overlap anything)")
    ap.add_argument("--no-long", action="store_true", help="skip the long_series object (configs[3])")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling run of configs[1]")
    ap.add_argument("--no-parity", action="store_true", help="long series: skip the parity block")
    ap.add_argument("--levels-out", default=None, help="write the per-level launch table (JSON) here")
    ap.add_argument("--workload", default="batch", choices=["batch", "long"],
                    help="batch = configs[1] headline + strong + long_series (default); long = only configs[3], printed as the line")
    ap.add_argument("--long-n", type=int, default=100_000_000)
    ap.add_argument("--long-ell", type=int, default=4)
    ap.add_argument("--long-dtype", default="float32")
    ap.add_argument("--parity-n", type=int, default=4_000_000)
    ap.add_argument("--sub", type=int, default=None, help="long workload: rows per sub-chunk (power of two)")
    ap.add_argument("--variant", type=int, default=0, help="force a kernel family (0 auto, 1 lane-per-row, 2 thread-per-node, 3 column-split, 4 warp-per-node DMMA)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    rk = Ranks()
    rank, world, dev, dist = rk.rank, rk.world, rk.dev, rk.dist
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    from cyclic_gps import _native, cyclic_reduction as cr
    _native.load()
    _native.VARIANT = args.variant

    if args.workload == "long":
        res = run_long(args, rk, standalone=True)
        if rank == 0:
            cpu = None
            if not args.no_cpu_baseline:
                ncpu = min(args.long_n, 1_000_000)
                dt_ = getattr(torch, args.long_dtype)
                v, dt = cpu_reference_rows_per_s(ncpu, args.long_ell, dt_, 3, warm=1)
                cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                       "sample": f"3 passes over a series of n={ncpu} (of {args.long_n}; host RAM bounds the reference at ~628 B/row), "
                                 f"l={args.long_ell}, {args.long_dtype}, {dt:.1f} s"}
            line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": res["dtype"], "data": "synthetic",
                    "config": {"workload": res["workload"], "n": res["n"], "ell": res["ell"], "sub_chunk_rows": res["sub_chunk_rows"],
                               "sub_chunks": res["sub_chunks"], "parallelism": res["parallelism"],
                               "l2": "inputs per step exceed L2; no explicit flush"},
                    "roofline": res["roofline"], "cpu_baseline": cpu, "e2e": None, "gpu_launches": res["gpu_launches"],
                    "clocks": res["clocks"], "parity": res["parity"], "loglik": res["loglik"]}
            emit(line)
        rk.close()
        return

    # the timed loops always run backward(): opt into the deferred non-PD report and the early release of the factors
    cr.EAGER_PD_CHECK = False
    cr.RELEASE_FACTORS_AFTER_BACKWARD = True
    B, n, ell = args.batch, args.n, args.ell
    dtype = getattr(torch, args.dtype)
    s = torch.empty((), dtype=dtype).element_size()
    R, O, x = make_inputs(dev, 1000 + rank, B, n, ell, dtype)
    Rr, Or, xr = R.requires_grad_(True), O.requires_grad_(True), x.requires_grad_(True)

    # ---- pass 1: the headline timing (weak scaling: B series on every rank; no per-launch events inside)
    clocks = ClockSampler(rk.local) if rank == 0 else None
    if clocks is not None:
        clocks.start()
    ms, step, checksum, timed_launches = time_batch(rk, cr, (Rr, Or, xr), args.steps, args.warmup)
    # ---- pass 2: same steps with an event pair around every launch (per-kernel durations, roofline)
    trace = LaunchTrace()
    _native.TRACE = trace
    trace.enabled = True
    for _ in range(args.steps):
        step()
    rk.sync()
    trace.enabled = False
    _native.TRACE = None
    keep_busy_until_sampled(clocks, step, rank, world, dist, dev)
    clk = clocks.stop() if clocks is not None else None
    launches_per_step = len(trace.records) // max(args.steps, 1)
    rows = B * n * world
    value = rows * args.steps / (ms * 1e-3)

    # ---- strong scaling of the same workload: B series in TOTAL, B / world per rank
    strong = None
    if not args.no_strong:
        if world == 1:
            strong = {"series_total": B, "series_per_gpu": B, "ms_per_step": ms / args.steps, "value": value, "unit": UNIT,
                      "note": "one GPU: identical to the headline run"}
            ms_g = time_graphed(rk, (R, O, x), args.steps, args.warmup)
            if ms_g is not None:
                strong["cuda_graph"] = {"ms_per_step": ms_g / args.steps, "value": B * n * args.steps / (ms_g * 1e-3), "unit": UNIT,
                                        "what": "the same step as ONE CUDA-graph replay (cyclic_gps.graphs.GraphedMahalAndDet)"}
        elif B % world == 0:
            Bs = B // world
            sub_t = tuple(t.detach()[:Bs].requires_grad_(True) for t in (R, O, x))
            ms_s, _, _, _ = time_batch(rk, cr, sub_t, args.steps, args.warmup)
            strong = {"series_total": B, "series_per_gpu": Bs, "ms_per_step": ms_s / args.steps,
                      "value": B * n * args.steps / (ms_s * 1e-3), "unit": UNIT}
            ms_g = time_graphed(rk, sub_t, args.steps, args.warmup)
            if ms_g is not None:
                strong["cuda_graph"] = {"ms_per_step": ms_g / args.steps, "value": B * n * args.steps / (ms_g * 1e-3), "unit": UNIT,
                                        "what": "the same step as ONE CUDA-graph replay (cyclic_gps.graphs.GraphedMahalAndDet)"}
            del sub_t

    # ---- end to end: HOST buffers in (time stamps + observations), per-series scalars out, copies inside the timed region.
    # The precision blocks are built ON the device from the time stamps (cyclic_gps.peg, the reference's
    # compute_posterior_precision), so 12 bytes per row cross PCIe instead of (2 l^2 + l) elements.
    e2e = train = None
    if not args.no_e2e:
        e2e, train = run_e2e(args, rk, cr, B, n, ell, dtype, R, O)

    # ---- configs[3]: the single long series over the same ranks
    del R, O, x, Rr, Or, xr, step
    torch.cuda.empty_cache()
    long_series = None
    if not args.no_long:
        long_series = run_long(args, rk, standalone=False)
        torch.cuda.empty_cache()

    if rank != 0:
        rk.close()
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
    from cyclic_gps import _native as _nat
    pks = _nat.tri_stride(dtype, ell)
    roofline["storage"] = (f"packed lower triangles inside the library ({pks} of {ell * ell} elements for D, the reduced diagonal blocks and Sigma_d of "
                           "the inner levels): real traffic is below the algorithmic bytes that `achieved` counts (SURVEY 8(d))") if pks else "full blocks"
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
            "config": batch_config(B, n, ell, args.dtype, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": timed_launches,
            "gpu_launches_per_step": timed_launches // max(args.steps, 1), "per_level_table_launches_per_step": launches_per_step, "clocks": clk,
            "loglik_checksum": checksum, "strong": strong, "train_step": train, "long_series": long_series}
    emit(line)
    rk.close()


if __name__ == "__main__":
    main()
