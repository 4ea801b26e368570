# ncu capture of the precision-builder kernels at configs[1] shapes (B = 128 series): run under gpurun
set -x
python tools/e2e_breakdown.py 128 > gpurun_out/plain_peg.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cr_peg_bwd -s 3 -c 1 -o gpurun_out/r2_peg_bwd -f python tools/e2e_breakdown.py 128 > gpurun_out/ncu_peg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_peg_fwd -s 3 -c 1 -o gpurun_out/r2_peg_fwd -f python tools/e2e_breakdown.py 128 > gpurun_out/ncu_peg2.log 2>&1
ls -la gpurun_out/r2_peg*.ncu-rep; tail -3 gpurun_out/ncu_peg.log
