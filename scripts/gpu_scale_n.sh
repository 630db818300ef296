# usage: bash scripts/gpu_scale_n.sh <N>   (under gpurun --gpus N): the default bench line on N GPUs of one box
N=$1
mkdir -p gpurun_out
timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --e2e-max-gb 4 > gpurun_out/r02_bench_default_${N}gpu.json 2> gpurun_out/r02_${N}gpu.err
tail -2 gpurun_out/r02_${N}gpu.err; cut -c1-300 gpurun_out/r02_bench_default_${N}gpu.json
