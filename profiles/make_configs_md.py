"""Renders profiles/r1_configs.md from profiles/r1_configs.json (written by bench_configs.py)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
d = json.load(open(os.path.join(HERE, "r1_configs.json")))
c1, c3, c5 = d["cfg1"], d["cfg3"], d["cfg5"]
f = lambda v, p=3: ("%%.%dg" % p) % v if v is not None else "—"
out = [f"""# Round 1 — the other BASELINE configs (`bench_configs.py`, one {d.get('gpu', 'B200')}, {d.get('host_cores', 16)} host cores)

GPU: CUDA events, median after 3 warm-ups.  Reference CPU path = `oracle/cr_oracle.py` (same ATen ops as the reference,
torch autograd for gradients), `torch.set_num_threads({d.get('host_cores', 16)})`, best of 2.  Raw numbers: `r1_configs.json`
(rendered by `make_configs_md.py`).  cfg2 / cfg4 are `bench.py` / `bench.py --workload long` (see `r1_summary.md`).

## cfg1 — n = 1000, l = 3, fp64: decompose + mahal_and_det + solve

| | ms |
|---|---|
| GPU (this repo) | {c1['gpu_ms']:.3f} |
| reference CPU path | {c1['reference_cpu_ms']:.2f} |
| dense Cholesky of the assembled 3000 x 3000 matrix (CPU) | {c1['dense_cholesky_cpu_ms']:.1f} |

Parity (max-abs-err / max-abs-ref): vs reference path mahal {c1['parity_vs_reference_path']['mahal']:.1e}, logdet {c1['parity_vs_reference_path']['logdet']:.1e},
solve {c1['parity_vs_reference_path']['solve']:.1e}; vs dense mahal {c1['parity_vs_dense']['mahal']:.1e}, logdet {c1['parity_vs_dense']['logdet']:.1e}, solve {c1['parity_vs_dense']['solve']:.1e}.
This size is host/latency-bound on the GPU: four sweeps of ten levels; the deep levels of each sweep run as one fused launch
(`cr_tpn_*_multi_kernel`) and the remaining time is mostly Python / launch overhead (`tools/small_n_profile.py`).

## cfg3 — CO2-shaped training step: n = 502 (240-unit gap), l = 16, fp64, obs_dim = 1, model resident on the GPU

| | GPU ms | reference CPU ms |
|---|---|---|
| log_likelihood forward + backward (two CR factorisations, hand-written backward) | {c3['gpu_train_step_ms']:.2f} | {c3['reference_cpu_train_step_ms']:.1f} |
| in-sample posterior (decompose + solve + inverse_blocks) | {c3['gpu_posterior_ms']:.2f} | {c3['reference_cpu_posterior_ms']:.1f} |

Parity: loglik {c3['parity']['loglik']:.1e}, parameter gradients {c3['parity']['param_grads']:.1e}, posterior mean {c3['parity']['posterior_mean']:.1e},
posterior covariance blocks {c3['parity']['posterior_cov_diag']:.1e} / {c3['parity']['posterior_cov_off']:.1e}.

## cfg5 — kalman_timing_script sweep: fp64, regular gaps, obs_dim = 2

posterior = decompose + solve + inverse_blocks (what `compute_insample_posterior` costs); loglik = `mahal_and_det` forward.
The numpy Kalman filter is the comparator of `cyclic_gps/kalman.py` (filterpy is not installed); its log-likelihood agrees with
the CR log-likelihood wherever both were run (column `worst parity` includes it).

| l | n | GPU posterior ms | GPU block-rows/s | GPU loglik ms | ref CPU posterior ms | ref CPU loglik ms | numpy Kalman loglik ms | worst parity |
|---|---|---|---|---|---|---|---|---|"""]
for r in c5:
    par = list(r.get("parity", {}).values()) + ([r["loglik_vs_kalman_rel"]] if "loglik_vs_kalman_rel" in r else [])
    out.append(f"| {r['l']} | {r['n']:.0e} | {r['gpu_posterior_ms']:.3f} | {r['gpu_posterior_rows_per_s']:.3g} | {r['gpu_loglik_ms']:.3f} | "
               f"{f(r.get('reference_cpu_posterior_ms'))} | {f(r.get('reference_cpu_loglik_ms'))} | {f(r.get('numpy_kalman_loglik_ms'))} | "
               f"{('%.1e' % max(par)) if par else '—'} |")
big8 = next((r for r in c5 if r["l"] == 8 and r["n"] == 10_000_000), None)
big32 = next((r for r in c5 if r["l"] == 32 and r["n"] >= 100_000), None)
out.append("")
if big8:
    gbs = 3264 * big8["n"] / (big8["gpu_loglik_ms"] * 1e-3) / 1e12
    out.append(f"Notes.  l = 2 runs on the thread-per-node kernels, l = 8 fp64 on the column-split kernels; at n = 1e7 the l = 8 fp64 log-likelihood pass moves "
               f"(6 l^2 + 3 l) * 8 B = 3264 B per row in {big8['gpu_loglik_ms']:.1f} ms = {gbs:.1f} TB/s ({100 * gbs / 6.5434:.0f} % of the measured HBM peak).")
if big32:
    gbs = (6 * 1024 + 96) * 8 * big32["n"] / (big32["gpu_loglik_ms"] * 1e-3) / 1e12
    out.append(f"l = 32 fp64 runs on the generic lane-per-row kernels (one warp per node): {gbs:.2f} TB/s, far from the roofline — the known weak spot, see DESIGN.md.")
out.append("Small n is host/launch-bound (see cfg1).")
open(os.path.join(HERE, "r1_configs.md"), "w").write("\n".join(out) + "\n")
print("wrote r1_configs.md")
