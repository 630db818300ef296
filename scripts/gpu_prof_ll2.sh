set -x
mkdir -p gpurun_out
S2="python bench.py --workload chol1024 --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_ll2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blk_ll2_kernel -s 11 -c 3 -f -o gpurun_out/prof_ll2 $S2 > gpurun_out/ncu_ll2.log 2>&1
tail -3 gpurun_out/ncu_ll2.log
S="python bench.py --workload chol1024 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S > gpurun_out/plain_c1024.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chol1024.csv $S > gpurun_out/ncu_c1024.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --workload chol1024 --cpu-seconds 8 --e2e-steps 1 > gpurun_out/r01_chol1024.json 2> gpurun_out/r01_chol1024.err; tail -2 gpurun_out/r01_chol1024.err
timeout 600 python -m pytest tests/test_gpu_tensor.py -m gpu -q 2>&1 | tail -2
