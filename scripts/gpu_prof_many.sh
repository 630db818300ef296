# usage: bash scripts/gpu_prof_many.sh "<tag> <kernel-regex> <kind> <n> <batch> [extra quick_time args]" ...
# one ncu --set full capture (third matching launch) per argument, each after the same command exited 0 without ncu
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec
  TAG=$1; RE=$2; shift 2
  CMD="python scripts/quick_time.py $*"
  timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$RE" -s 2 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
  cat gpurun_out/plain_$TAG.log; tail -1 gpurun_out/ncu_$TAG.log
  # the reports are large (source import): keep the text summary, drop the report unless KEEP_REP=1
  python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep "scripts/gpu_prof_many.sh: quick_time.py $*" > gpurun_out/summary_$TAG.txt 2>&1
  [ "$KEEP_REP" = "1" ] || rm -f gpurun_out/prof_$TAG.ncu-rep
done
