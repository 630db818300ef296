set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest2.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1"
for mode in 0 1; do for w in 4 5 6; do HLSD_TPM_MODE=$mode HLSD_TPM_WARPS=$w timeout 200 $B > gpurun_out/b_mode${mode}_w${w}.json 2>gpurun_out/b_err.log; done; done
for wl in lu64 qr128 chol64; do timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl --cpu-seconds 5 --e2e-steps 1 > gpurun_out/b_$wl.json 2> gpurun_out/b_$wl.err; done
S="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 200 $S > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chol16.csv $S > gpurun_out/ncu_launch.log 2>&1
timeout 200 $S > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chol_tpm -s 3 -c 2 -o gpurun_out/prof_chol16 $S > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest2.log
for f in gpurun_out/b_mode*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_median'])"; done
