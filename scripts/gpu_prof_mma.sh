set -x
mkdir -p gpurun_out
S="python bench.py --workload chol1024 --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S > gpurun_out/plain_mma.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blk_mma_kernel -s 30 -c 4 -f -o gpurun_out/prof_mma $S > gpurun_out/ncu_mma.log 2>&1
tail -3 gpurun_out/ncu_mma.log
