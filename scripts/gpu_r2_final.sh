# Round-2 evidence refresh, part A (one GPU): full GPU suite, smoke, default bench (headline + workloads + per-N), the
# reference arm, launch lists of the two N = 1024 workloads.  Part B (ncu --set full captures, summarised on the box
# because the reports exceed what gpurun copies back): scripts/gpu_r2_final_ncu.sh
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_gputests.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee -a gpurun_out/r02_gputests.txt
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -2 gpurun_out/r02_bench_default.err; cut -c1-300 gpurun_out/r02_bench_default.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
for w in lu1024 chol1024; do
  S="python bench.py --workload $w --batch 592 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --headline-only"
  timeout 300 $S > gpurun_out/plain_$w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_${w}_launches.csv $S > gpurun_out/ncu_$w.log 2>&1
done
