# ncu evidence of the binary with packed lower triangles (run under gpurun, one GPU).  Every profiled command first exits 0 WITHOUT ncu.
set -x
H="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-long --no-strong"
$H > gpurun_out/r2_tri_plain_headline.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cr_ -c 400 --csv --log-file gpurun_out/r2_tri_launches_ncu.csv $H > gpurun_out/r2_tri_ncu_launches.log 2>&1
B5="python bench.py --batch 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-long --no-strong"
$B5 > gpurun_out/r2_tri_plain_b512.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cr_tpn_fwd_kernel -s 24 -c 1 -o gpurun_out/r2_tri_tpn_fwd_f32_8 -f $B5 > gpurun_out/r2_tri_ncu_tpn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_tpn_bwd_kernel -s 31 -c 1 -o gpurun_out/r2_tri_tpn_bwd_f32_8 -f $B5 > gpurun_out/r2_tri_ncu_tpn_bwd.log 2>&1
# a level-1 launch of each direction (inner level: packed in AND out)
ncu --set full --clock-control none --import-source on -k regex:cr_tpn_fwd_kernel -s 25 -c 1 -o gpurun_out/r2_tri_tpn_fwd_f32_8_l1 -f $B5 > gpurun_out/r2_tri_ncu_tpn_fwd1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_tpn_bwd_kernel -s 30 -c 1 -o gpurun_out/r2_tri_tpn_bwd_f32_8_l1 -f $B5 > gpurun_out/r2_tri_ncu_tpn_bwd1.log 2>&1
for r in r2_tri_tpn_fwd_f32_8 r2_tri_tpn_bwd_f32_8 r2_tri_tpn_fwd_f32_8_l1 r2_tri_tpn_bwd_f32_8_l1; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  python tools/ncu_src_summary.py gpurun_out/$r.ncu-rep > gpurun_out/${r}_ncu_summary.txt 2>&1
done
rm -f gpurun_out/r2_tri_tpn_fwd_f32_8_l1.ncu-rep gpurun_out/r2_tri_tpn_bwd_f32_8_l1.ncu-rep gpurun_out/r2_tri_tpn_fwd_f32_8.ncu-rep   # (merge limit: 64 MiB per call)
ls -la gpurun_out/
