set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --workload chol64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_chol64.json 2> gpurun_out/b_chol64.err; cat gpurun_out/b_chol64.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/b_chol16_e2e.json 2> gpurun_out/b_chol16.err; cat gpurun_out/b_chol16_e2e.json
