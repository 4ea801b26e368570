B32="python bench.py --ell 32 --dtype float64 --batch 24 --n 10000 --steps 1 --warmup 3 --no-long --no-strong --no-cpu-baseline --no-e2e"
$B32 --levels-out gpurun_out/levels_l32_tw2.json > gpurun_out/plain32.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cr_mma_bwd -s 13 -c 1 -o gpurun_out/r2_mma_bwd_f64_32_tw2 -f $B32 > gpurun_out/ncu2.log 2>&1
