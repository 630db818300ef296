set -x
mkdir -p gpurun_out
export HLSD_LL_TS=1
S2="python bench.py --workload chol1024 --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_llts.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blk_ll_ts_kernel -s 12 -c 2 -f -o gpurun_out/prof_llts $S2 > gpurun_out/ncu_llts.log 2>&1
tail -3 gpurun_out/ncu_llts.log
