#!/bin/bash
# compute-sanitizer over the backward kernels and the row-trimmed multi-part path.
mkdir -p gpurun_out
SEL='backward or autograd or row_trimmed'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer2_$tool.log | tail -3
done
