# round 2, call 1: box facts, GPU test suite, default bench (all workloads), baseline ncu captures of the
# kernels that had no committed summary in round 1
mkdir -p gpurun_out
{ nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; nvidia-smi topo -m; nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm --format=csv; } > gpurun_out/sysinfo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_default.err
prof() {  # <tag> <kernel regex> <kind> <n> <batch>
  CMD="python scripts/quick_time.py $3 $4 $5"
  timeout 300 $CMD > gpurun_out/plain_$1.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  cat gpurun_out/plain_$1.log; tail -1 gpurun_out/ncu_$1.log
}
prof lu64 lu_lane_kernel lu 64 65536
prof qr128 qr_qreg_kernel qr 128 2048
prof lupanel lu_panel_kernel lu 1024 256
prof lugemm "gemm_tile_kernel<1>" lu 1024 256
prof choldiag "blk_diag_kernel<0" chol 1024 256
