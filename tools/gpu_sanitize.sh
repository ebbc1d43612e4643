#!/bin/bash
# compute-sanitizer over a small but representative subset of the GPU tests.
mkdir -p gpurun_out
SEL='test_all_tma_kernel_every_variant or test_plane_generic_nhwc_identical_bits or test_cur_frame_path_as_shipped or test_fused_golden_fixture or test_req_add_and_null or test_gpu_mv_accumulate_bit_exact or test_coviar_front_end or test_all_tma_static_and_dynamic'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
done
