set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py -m gpu -q -x -s 2>&1 | tail -12
S="python bench.py --workload lu1024 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S > gpurun_out/b_lu1024_b512.json 2> gpurun_out/b_lu1024.err; cut -c1-200 gpurun_out/b_lu1024_b512.json; tail -3 gpurun_out/b_lu1024.err
timeout 300 $S > gpurun_out/plain_lu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lu1024.csv $S > gpurun_out/ncu_lu.log 2>&1
timeout 600 python bench.py --workload lu1024 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_lu1024_full.json 2>> gpurun_out/b_lu1024.err; cut -c1-200 gpurun_out/b_lu1024_full.json
