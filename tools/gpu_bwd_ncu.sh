#!/bin/bash
# ncu of the backward kernels.  Usage: bash tools/gpu_bwd_ncu.sh tag
TAG=${1:-b}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/bwd_launches_$TAG.csv \
    python tools/bench_configs.py --only-backward --quick > /dev/null 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_nchw_gather -s 2 -c 2 -f -o gpurun_out/prof_bwd_$TAG \
    python tools/bench_configs.py --only-backward > gpurun_out/ncu_bwd_$TAG.log 2>&1; echo "full rc=$?"
