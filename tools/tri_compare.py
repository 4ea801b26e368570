"""Packed lower triangles vs full blocks: prints the two bench lines and the per-level launch times written by
`CRB200_TRI=1|0 python bench.py --no-long --no-strong --no-cpu-baseline --no-e2e --levels-out gpurun_out/s3_levels_tri$t.json > gpurun_out/s3_bench_tri$t.json`."""
import json,sys
for t in (1,0):
    d=json.loads(open(f"gpurun_out/s3_bench_tri{t}.json").read().strip().splitlines()[-1])
    print(t, d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["whole_step"]["frac"], d["clocks"])
a=json.load(open('gpurun_out/s3_levels_tri1.json')); b=json.load(open('gpurun_out/s3_levels_tri0.json'))
for x,y in list(zip(a,b))[:12]:
    print(x.get('kind'), x.get('m'), round(x.get('avg_ms'),4), round(y.get('avg_ms'),4), round(x['avg_ms']/y['avg_ms'],3))
