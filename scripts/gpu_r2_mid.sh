# Mid-round evidence refresh: default bench (headline + workloads + per-N), LU 1024 launch list, ncu summaries of the LU kernels
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -2 gpurun_out/bench_r2b.err; cut -c1-400 gpurun_out/bench_r2b.json
S="python bench.py --workload lu1024 --batch 592 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --headline-only"
timeout 300 $S > gpurun_out/plain_lu1024b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lu1024_crout.csv $S > gpurun_out/ncu_lu1024b.log 2>&1
tail -2 gpurun_out/plain_lu1024b.log | cut -c1-300
