# Second evidence refresh of the round: tests, sweep, the non-headline bench lines, ncu captures of the LU 16 and QR 16 kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in lu64 lu64np qr128 chol64 lu1024; do timeout 900 python bench.py --steps 5 --warmup 3 --workload $wl --cpu-seconds 8 --e2e-steps 1 > gpurun_out/r01_$wl.json 2> gpurun_out/r01_$wl.err; tail -2 gpurun_out/r01_$wl.err; done
timeout 1500 python scripts/sweep.py > gpurun_out/r01_sweep.md 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.err
bash scripts/gpu_prof_kernel.sh lu_regs_kernel lu 16 1027176 > /dev/null 2>&1
bash scripts/gpu_prof_kernel.sh qr_wcols_kernel qr 16 1048576 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
