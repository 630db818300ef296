mkdir -p gpurun_out
cap() {
  K=$1; S=$2; C=$3; TAG=$4; shift 4
  CMD="python scripts/quick_time.py $*"
  timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { cat gpurun_out/plain_$TAG.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
  { python scripts/ncu_summary.py /tmp/prof_$TAG.ncu-rep "quick_time.py $*, kernel $K, launch $S (+$C)"; echo; echo "# per source line (scripts/ncu_lines.py)"; python scripts/ncu_lines.py /tmp/prof_$TAG.ncu-rep 30; } > gpurun_out/r02_${TAG}_ncu_summary.txt 2>&1
  rm -f /tmp/prof_$TAG.ncu-rep
  head -2 gpurun_out/plain_$TAG.log; grep "gpu__time_duration" gpurun_out/r02_${TAG}_ncu_summary.txt | head -1
}
cap qr_qreg 1 1 qr128_final qr 128 2048
# (the cholesky_v2.2 kernels are reached with --design; quick_time.py times the default design only)
cap blk_diag2 3 1 blk_diag2_final chol 1024 592
