"""loglik+grad throughput across block sizes and dtypes (bench.py's batch workload at a size that fits quickly):
writes a JSON list with ms/step, block-rows/s and the whole-step fraction of the measured HBM peak.
usage: python tools/ell_sweep.py out.json [ell,ell,...] [variant]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
ells = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 24, 32]
variant = sys.argv[3] if len(sys.argv) > 3 else "0"
cases = [(l, dt) for dt in ("float32", "float64") for l in ells]
for l, dt in cases:
    s = 4 if dt == "float32" else 8
    rows = int(min(5.12e6, 2.0e9 / (l * l * s)))            # keep inputs + factors around a few GB
    batch = max(8, rows // 10000)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--ell", str(l), "--dtype", dt, "--batch", str(batch), "--n", "10000", "--no-long", "--no-strong",
           "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--variant", variant]
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        out.append({"ell": l, "dtype": dt, "error": r.stderr[-300:]})
        continue
    out.append({"ell": l, "dtype": dt, "variant": int(variant), "batch": batch, "n": 10000, "ms_per_step": d["ms_per_step"], "rows_per_s": d["value"],
                "whole_step_frac_of_hbm_peak": d["roofline"]["whole_step"]["frac"], "top_kernel": d["roofline"]["kernel"],
                "top_kernel_frac": d["roofline"]["frac"], "gpu_launches_per_step": d.get("gpu_launches_per_step")})
    print(out[-1], flush=True)
json.dump(out, open(sys.argv[1], "w"), indent=1)
