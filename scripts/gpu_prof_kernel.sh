# usage: bash scripts/gpu_prof_kernel.sh <kernel-regex> <chol|lu|qr> <n> <batch> [tag]
# full ncu capture of one launch of the named kernel (after the same command exited 0 without ncu)
mkdir -p gpurun_out
TAG=${5:-$2$3}
CMD="python scripts/quick_time.py $2 $3 $4"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/plain_$TAG.log; tail -2 gpurun_out/ncu_$TAG.log
