# whole-box scaling check: the four BASELINE.json configurations on all 8 GPUs (weak scaling, batch split by rank)
mkdir -p gpurun_out
N=${1:-8}   # usage: bash scripts/gpu_scale8.sh [gpus] ["workload ..."]
for wl in ${2:-chol16 lu64 qr128 chol1024}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01_${wl}_${N}gpu.json 2> gpurun_out/r01_${wl}_${N}gpu.err
  tail -1 gpurun_out/r01_${wl}_${N}gpu.json | cut -c1-220
done
