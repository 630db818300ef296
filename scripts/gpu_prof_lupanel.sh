set -x
mkdir -p gpurun_out
S2="python bench.py --workload lu1024 --batch 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_lupanel.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lu_panel_kernel -s 24 -c 2 -f -o gpurun_out/prof_lupanel $S2 > gpurun_out/ncu_lupanel.log 2>&1
tail -3 gpurun_out/ncu_lupanel.log
