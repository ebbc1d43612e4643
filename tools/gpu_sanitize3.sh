#!/bin/bash
# compute-sanitizer over the kernels added in the third session: channels-last all-TMA kernel (record warp, 2 producer
# warps, 2 consumer groups), residual front end, NCHW warp-only variant with direct stores, the MV row copy helper.
mkdir -p gpurun_out
SEL='nhwc_all_tma_kernel_every_variant or nhwc_all_tma_kernel_random or res_coviar or mv_centre_rows or host_aggregator or test_bilinear_sampler or cur_frame_path'
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer3_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer3_$tool.log | tail -3
done
