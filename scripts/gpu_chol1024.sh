set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor.py -m gpu -q -x 2>&1 | tail -3
S="python bench.py --workload chol1024 --batch 512 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
timeout 300 $S > gpurun_out/b_chol1024_b512.json 2> gpurun_out/b_chol1024.err
cat gpurun_out/b_chol1024_b512.json; tail -3 gpurun_out/b_chol1024.err
timeout 300 $S > gpurun_out/plain5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chol1024.csv $S > gpurun_out/ncu5.log 2>&1
timeout 600 python bench.py --workload chol1024 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_chol1024_full.json 2>> gpurun_out/b_chol1024.err
cat gpurun_out/b_chol1024_full.json
