set -x
mkdir -p gpurun_out
S2="python bench.py --workload chol1024 --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_ll.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blk_ll_kernel -s 11 -c 3 -f -o gpurun_out/prof_ll $S2 > gpurun_out/ncu_ll.log 2>&1
tail -3 gpurun_out/ncu_ll.log
