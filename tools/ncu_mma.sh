set -x
B32="python bench.py --ell 32 --dtype float64 --batch 24 --n 10000 --steps 1 --warmup 3 --no-long --no-strong --no-cpu-baseline --no-e2e"
B16="python bench.py --ell 16 --dtype float64 --batch 97 --n 10000 --steps 1 --warmup 3 --no-long --no-strong --no-cpu-baseline --no-e2e"
$B32 > gpurun_out/plain32.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cr_mma_fwd -c 1 -o gpurun_out/r2_mma_fwd_f64_32 -f $B32 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_mma_bwd -s 13 -c 1 -o gpurun_out/r2_mma_bwd_f64_32 -f $B32 > gpurun_out/ncu2.log 2>&1
$B16 > gpurun_out/plain16.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cr_mma_fwd -c 1 -o gpurun_out/r2_mma_fwd_f64_16 -f $B16 > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cr_mma_bwd -s 13 -c 1 -o gpurun_out/r2_mma_bwd_f64_16 -f $B16 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu1.log gpurun_out/ncu4.log
