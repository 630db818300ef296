set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "qr" 2>&1 | tail -4
timeout 300 python bench.py --workload qr128 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/b_qr128.json 2> gpurun_out/b_qr128.err; cat gpurun_out/b_qr128.json; tail -2 gpurun_out/b_qr128.err
