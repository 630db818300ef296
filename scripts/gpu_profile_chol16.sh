set -x
mkdir -p gpurun_out
S="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 200 $S > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chol16.csv $S > gpurun_out/ncu_launch.log 2>&1
timeout 200 $S > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chol_rows -s 3 -c 2 -f -o gpurun_out/prof_chol16 $S > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_chol16_full.json 2> gpurun_out/bench_full.err
cat gpurun_out/bench_chol16_full.json
