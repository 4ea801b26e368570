"""Summarise the source page of an ncu report: opcode histogram weighted by stall samples / executed instructions, stall reasons,
and the hottest source lines.  usage: python tools/ncu_src_summary.py report.ncu-rep [kernel-source-file-substring]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
if len(rr) >= 3:
    d = dict(zip(rr[0], rr[2]))
    for k in ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
              "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
              "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
        if k in d:
            print(f"{k:75s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
data = rows[hi + 1:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("samples", tot, "warp instructions", inst, "SASS lines", len(data))
cs, ci = Counter(), Counter()
for r in data:
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    cs[op] += int(r[ix["# Samples"]])
    ci[op] += int(r[ix["Instructions Executed"]])
for op, c in cs.most_common(18):
    print(f"{op:10s} samples {100 * c / tot:5.1f}%  instructions {100 * ci[op] / inst:5.1f}%")
st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
sc = {k: sum(int(r[ix[k]]) for r in data) for k in st}
print(" ".join(f"{k[6:]}={100 * v / tot:.1f}%" for k, v in sorted(sc.items(), key=lambda x: -x[1])[:9]))
