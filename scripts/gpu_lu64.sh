mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lu or edge or device" 2>&1 | tail -4
for mode in 0 1; do HLSD_LU_ROWS=$mode timeout 300 python bench.py --workload lu64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/lu64_m$mode.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/lu64_m$mode.json')); print('mode', $mode, d['value'], d['roofline']['frac'], d['ms_per_step'])"; done
