mkdir -p gpurun_out
S2="python bench.py --workload qr128 --batch 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_qr.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qr_wave_kernel -s 2 -c 1 -f -o gpurun_out/prof_qr $S2 > gpurun_out/ncu_qr.log 2>&1
tail -2 gpurun_out/ncu_qr.log
