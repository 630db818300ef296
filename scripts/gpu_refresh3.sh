# Final evidence refresh of the round: full GPU suite, smoke, LU 1024 bench line + launch list, size sweep.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --workload lu1024 --cpu-seconds 8 --e2e-steps 1 > gpurun_out/r01_lu1024.json 2> gpurun_out/r01_lu1024.err; tail -2 gpurun_out/r01_lu1024.err
S="python bench.py --workload lu1024 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S > gpurun_out/plain_lu1024.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lu1024.csv $S > gpurun_out/ncu_lu1024.log 2>&1
timeout 1500 python scripts/sweep.py > gpurun_out/r01_sweep.md 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.err
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
