#!/bin/bash
# backward ops: parity tests + kernel table.  Usage: bash tools/gpu_bwd.sh tag
TAG=${1:-b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or autograd" 2>&1 | tail -25
timeout 300 python tools/bench_configs.py --only-backward > gpurun_out/bwd_$TAG.jsonl 2> gpurun_out/bwd_$TAG.err; echo "rc=$?"
cat gpurun_out/bwd_$TAG.jsonl; tail -5 gpurun_out/bwd_$TAG.err
