"""Opcode histogram (weighted by executed warp-instructions) and stall share from an
`ncu -i X.ncu-rep --page source --csv --print-source sass` dump.
usage: sass_hist.py file.csv [kernel-index]"""
import collections
import csv
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
kernels, cur, hdr = [], None, None
for row in csv.reader(open(path)):
    if len(row) >= 2 and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        kernels.append(cur)
        hdr = None
        continue
    if cur is None:
        continue
    if hdr is None:
        hdr = row
        cur["hdr"] = hdr
        continue
    cur["rows"].append(row)
k = kernels[which]
h = {n: i for i, n in enumerate(k["hdr"])}
ops, stalls, wf, wfi = (collections.Counter() for _ in range(4))
tot = samples = 0
for r in k["rows"]:
    try:
        n = int(r[h["Instructions Executed"]])
        s = int(r[h["# Samples"]])
    except Exception:
        continue
    toks = r[h["Source"]].strip().split()
    op = (toks[0] if not toks[0].startswith("@") else toks[1]).rstrip(";")
    ops[op] += n
    tot += n
    stalls[op] += s
    samples += s
    try:
        wf[op] += int(r[h["L1 Wavefronts Shared"]])
        wfi[op] += int(r[h["L1 Wavefronts Shared Ideal"]])
    except Exception:
        pass
print(k["name"], "| kernels in file:", len(kernels), "| warp-instr:", tot, "| samples:", samples)
for op, n in ops.most_common(30):
    print(f"  {op:28s} {n:12d} {100*n/tot:5.1f}%  stall {100*stalls[op]/max(samples,1):5.1f}%  smem wavefronts {wf[op]} (ideal {wfi[op]})")
