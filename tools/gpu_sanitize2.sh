#!/bin/bash
# compute-sanitizer over the kernels added / changed in the second session: backward (lists, gather, overflow,
# scatter), row-trimmed multi-part loads, hybrid work split, all-TMA cosine pre-pass, NHWC prefetch.
mkdir -p gpurun_out
SEL='backward or autograd or row_trimmed or cosine_prepass or test_all_tma_kernel_every_variant or test_cosine_logits_op or test_fused_golden_fixture'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer2_$tool.log | tail -3
done
