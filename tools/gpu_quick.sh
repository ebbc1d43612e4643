#!/bin/bash
# Quick GPU visit: parity tests + bench (+ optional ncu of the plane kernel).  Usage: bash tools/gpu_quick.sh tag [ncu]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nchw_tma -s 3 -c 1 -f -o gpurun_out/prof_plane_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
