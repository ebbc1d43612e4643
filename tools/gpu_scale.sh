#!/bin/bash
# Multi-GPU bench exactly as the driver launches it.  Usage: bash tools/gpu_scale.sh N tag
N=${1:-2}; TAG=${2:-s}
mkdir -p gpurun_out
if [ "$N" == "1" ]; then
  python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
fi
echo "rc=$?"; tail -c 1500 gpurun_out/scale_${TAG}_n$N.json; tail -5 gpurun_out/scale_${TAG}_n$N.err
