# Round evidence refresh: full GPU test suite, smoke, bench lines, launch list + full ncu capture of the
# headline kernel, size sweep.  Everything lands in gpurun_out/ (copied into profiles/ afterwards).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
S="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 200 $S > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chol16.csv $S > gpurun_out/ncu_launch.log 2>&1
timeout 200 $S > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:chol_rows -s 3 -c 2 -f -o gpurun_out/prof_chol16 $S > gpurun_out/ncu_full.log 2>&1
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/bench_ref_default.json 2> gpurun_out/bench_ref_default.err; tail -4 gpurun_out/bench_ref_default.err
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_chol16_full.json 2> gpurun_out/bench_full.err
cat gpurun_out/bench_chol16_full.json
for wl in lu64 qr128 chol64 chol1024 lu1024; do timeout 900 python bench.py --steps 5 --warmup 3 --workload $wl --cpu-seconds 8 --e2e-steps 1 > gpurun_out/r01_$wl.json 2> gpurun_out/r01_$wl.err; tail -2 gpurun_out/r01_$wl.err; done
timeout 1500 python scripts/sweep.py > gpurun_out/r01_sweep.md 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.err
