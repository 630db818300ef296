set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1"
for mode in tpm r2 r4; do
  HLSD_CHOL16=$mode timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cholesky or edge or device" 2>&1 | tail -4 > gpurun_out/pytest3_$mode.log
  HLSD_CHOL16=$mode timeout 200 $B > gpurun_out/c_$mode.json 2>gpurun_out/c_err.log
done
for f in gpurun_out/c_*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_median'])"; done
cat gpurun_out/pytest3_*.log
