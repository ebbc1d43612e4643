#!/bin/bash
# Round-2 GPU visit A: bring-up of the tcgen05 convolutions + the new reference pins.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/tc_debug.py > gpurun_out/r2a_tc_debug.log 2>&1; echo "tc_debug rc=$?"; cat gpurun_out/r2a_tc_debug.log | tail -60
timeout 600 python -m pytest tests/test_tc_convs.py -q 2>&1 | tail -25
timeout 600 python -m pytest tests/test_cudnn_pin.py tests/test_coviar_ref.py -q 2>&1 | tail -40
( time timeout 900 python -m pytest tests -m gpu -q --ignore=tests/test_tc_convs.py --ignore=tests/test_cudnn_pin.py --ignore=tests/test_coviar_ref.py 2>&1 | tail -8 ) 2>&1
timeout 120 python tools/make_golden_cudnn.py 2>&1 | tail -3
timeout 400 python tools/bench_configs.py --only-keyframe > gpurun_out/r2a_keyframe.jsonl 2> gpurun_out/r2a_keyframe.err; echo "keyframe rc=$?"; cut -c1-260 gpurun_out/r2a_keyframe.jsonl; tail -5 gpurun_out/r2a_keyframe.err
