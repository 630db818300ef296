set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lu or edge or device" 2>&1 | tail -4
timeout 300 python bench.py --workload lu64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_lu64.json 2> gpurun_out/b_lu64.err; cat gpurun_out/b_lu64.json; tail -2 gpurun_out/b_lu64.err
