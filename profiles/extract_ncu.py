"""Selected metrics of `ncu -i X.ncu-rep --page raw --csv` dumps -> one JSON (units kept).
usage: extract_ncu.py out.json raw1.csv [raw2.csv ...]"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"] + [
        "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in
        ("long_scoreboard", "no_instruction", "wait", "short_scoreboard", "not_selected", "dispatch_stall",
         "lg_throttle", "mio_throttle", "branch_resolving")]

out = {}
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    for r in rows[2:]:
        name = "%s@grid%s" % (r[col["Kernel Name"]].replace("void ", "").split("(")[0], r[col["Grid Size"]].strip("()").split(",")[0])
        ent = {}
        for k in KEEP:
            if k in col:
                try:
                    ent[k] = [float(r[col[k]].replace(",", "")), units[col[k]]]
                except ValueError:
                    ent[k] = [r[col[k]], units[col[k]]]
        out[name] = ent
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, v in out.items():
    rd, wr = v["dram__bytes_read.sum"], v["dram__bytes_write.sum"]
    print(k, v["gpu__time_duration.sum"], "dram", rd, wr)
