set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/bench_ref_default.json 2> gpurun_out/bench_ref_default.err; tail -4 gpurun_out/bench_ref_default.err; cut -c1-400 gpurun_out/bench_ref_default.json
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
for wl in lu64 qr128 chol64 chol1024; do timeout 900 python bench.py --steps 5 --warmup 3 --workload $wl --cpu-seconds 8 --e2e-steps 1 > gpurun_out/r01_$wl.json 2> gpurun_out/r01_$wl.err; tail -2 gpurun_out/r01_$wl.err; done
S="python bench.py --workload chol1024 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S > gpurun_out/plain_c1024.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chol1024.csv $S > gpurun_out/ncu_c1024.log 2>&1
S2="python bench.py --workload chol1024 --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $S2 > gpurun_out/plain_mma.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blk_mma_kernel -s 30 -c 4 -f -o gpurun_out/prof_mma $S2 > gpurun_out/ncu_mma.log 2>&1
