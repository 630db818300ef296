mkdir -p gpurun_out
S2="python bench.py --workload chol1024 --batch 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_diag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blk_diag_kernel -s 9 -c 1 -f -o gpurun_out/prof_diag $S2 > gpurun_out/ncu_diag.log 2>&1
tail -2 gpurun_out/ncu_diag.log
