#!/bin/bash
# config 5: multi-stream sweep on 1 GPU and (optionally) N GPUs.  Usage: bash tools/gpu_streams.sh tag [N]
TAG=${1:-s}; N=${2:-1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "accumulate or coviar or stream_scheduler" 2>&1 | tail -2
python tools/bench_configs.py --only-upstream 2>&1 | tail -1
timeout 600 python tools/bench_streams.py > gpurun_out/streams_${TAG}_n1.jsonl 2> gpurun_out/streams_${TAG}_n1.err; echo "rc=$?"; cat gpurun_out/streams_${TAG}_n1.jsonl; tail -3 gpurun_out/streams_${TAG}_n1.err
if [ "$N" != "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/bench_streams.py > gpurun_out/streams_${TAG}_n$N.jsonl 2> gpurun_out/streams_${TAG}_n$N.err; echo "rc=$?"; cat gpurun_out/streams_${TAG}_n$N.jsonl; tail -3 gpurun_out/streams_${TAG}_n$N.err
  bash tools/gpu_scale.sh $N $TAG
fi
