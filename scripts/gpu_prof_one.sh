# usage: bash scripts/gpu_prof_one.sh <kernel-regex> <skip> <count> <tag> <quick_time args...>
# full ncu capture of launches [skip, skip+count) of the named kernel (after the same command exited 0 without ncu)
mkdir -p gpurun_out
K=$1; S=$2; C=$3; TAG=$4; shift 4
CMD="python scripts/quick_time.py $*"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { cat gpurun_out/plain_$TAG.log; exit 1; }
cat gpurun_out/plain_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1; tail -1 gpurun_out/ncu_$TAG.log
