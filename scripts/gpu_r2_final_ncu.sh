# Round-2 evidence refresh, part B: ncu --set full captures of the tensor-LU kernels, the LU lane kernels and the Cholesky
# update kernel, each summarised on the box (scripts/ncu_summary.py + per-source-line stalls) and the report deleted.
mkdir -p gpurun_out
cap() {  # cap <kernel-regex> <skip> <count> <tag> <quick_time args...>
  K=$1; S=$2; C=$3; TAG=$4; shift 4
  CMD="python scripts/quick_time.py $*"
  timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { cat gpurun_out/plain_$TAG.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
  { python scripts/ncu_summary.py /tmp/prof_$TAG.ncu-rep "scripts/gpu_r2_final_ncu.sh: quick_time.py $*, kernel $K, launch $S (+$C)"; echo; echo "# per source line (scripts/ncu_lines.py)"; python scripts/ncu_lines.py /tmp/prof_$TAG.ncu-rep 40; } > gpurun_out/r02_${TAG}_ncu_summary.txt 2>&1
  rm -f /tmp/prof_$TAG.ncu-rep
  head -3 gpurun_out/plain_$TAG.log; grep "gpu__time_duration" gpurun_out/r02_${TAG}_ncu_summary.txt | head -2
}
cap lu_cpanel 0 1 lu_cpanel lu 1024 296
cap lu_ll_kernel 7 2 lu_ll lu 1024 296
cap lu_assemble 0 1 lu_assemble lu 1024 296
cap lu_trsm 2 1 lu_trsm lu 1024 296
cap lu_lane 1 1 lu32_final lu 32
cap lu_lane 1 1 lu64_final lu 64
cap lu_lane 1 1 lu128_final lu 128
cap blk_ll2 4 1 ll2_final chol 1024 592
