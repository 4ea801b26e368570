# Round-2 ncu evidence of the final binary (run under gpurun, one GPU).  Every profiled command first exits 0 WITHOUT ncu.
set -x
H="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-long --no-strong"
$H > gpurun_out/r2_plain_headline.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cr_ -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv $H > gpurun_out/r2_ncu_launches.log 2>&1
B5="python bench.py --batch 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-long --no-strong"
$B5 > gpurun_out/r2_plain_b512.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cr_tpn_fwd_kernel -s 24 -c 1 -o gpurun_out/r2_tpn_fwd_f32_8 -f $B5 > gpurun_out/r2_ncu_tpn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_tpn_bwd_kernel -s 31 -c 1 -o gpurun_out/r2_tpn_bwd_f32_8 -f $B5 > gpurun_out/r2_ncu_tpn_bwd.log 2>&1
P="python tools/e2e_breakdown.py 128"
$P > gpurun_out/plain_peg.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cr_peg_bwd -s 3 -c 1 -o gpurun_out/r2_peg_bwd -f $P > gpurun_out/ncu_peg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_peg_fwd -s 3 -c 1 -o gpurun_out/r2_peg_fwd -f $P > gpurun_out/ncu_peg2.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep gpurun_out/r2_launches_ncu.csv
