#!/bin/bash
# One script for every GPU visit (run under gpurun from the repo root):
#   gpurun --timeout 1800 -- 'bash tools/gpu_visit.sh tests'      full -m gpu suite
#   ... tools/gpu_visit.sh bench [tag]       bench.py (both arms) -> gpurun_out/bench_<tag>.json
#   ... tools/gpu_visit.sh tables [tag]      tools/bench_configs.py + bench_streams.py -> gpurun_out/*_<tag>.jsonl
#   ... tools/gpu_visit.sh profile [tag]     ncu launch list + --set full captures of the main kernels
#   ... tools/gpu_visit.sh sanitize [tag]    compute-sanitizer memcheck / synccheck / racecheck over a test selection
WHAT=${1:-tests}; TAG=${2:-r2}
mkdir -p gpurun_out
case "$WHAT" in
tests)
  ( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 ) 2>&1 | tail -20 ;;
bench)
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
  SMI=$!
  python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$?"
  python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
  kill $SMI
  python __graft_entry__.py --smoke 2>&1 | tail -2
  tail -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err ;;
tables)
  timeout 1200 python tools/bench_configs.py > gpurun_out/configs_table_$TAG.jsonl 2> gpurun_out/configs_table_$TAG.err; echo "configs rc=$?"
  timeout 600 python tools/bench_streams.py --streams 64 256 1024 > gpurun_out/streams_sweep_$TAG.jsonl 2>&1; echo "streams rc=$?"
  timeout 600 python tools/bench_configs.py --only-motion > gpurun_out/motion_distributions_$TAG.jsonl 2>&1; echo "motion rc=$?"
  timeout 600 python tools/bench_configs.py --only-keyframe > gpurun_out/keyframe_graphs_$TAG.jsonl 2>&1; echo "keyframe rc=$?"
  timeout 600 python tools/bench_configs.py --only-nhwc-win > gpurun_out/nhwc_window_$TAG.jsonl 2>&1; echo "window rc=$?"
  timeout 600 python tools/bench_configs.py --only-single > gpurun_out/single_frame_$TAG.jsonl 2>&1; echo "single rc=$?"
  cut -c1-200 gpurun_out/configs_table_$TAG.jsonl ;;
profile)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nchw_tma -s 3 -c 1 -f -o gpurun_out/prof_tma_$TAG \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > /dev/null 2>&1; echo "ncu headline rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 4 -c 4 -f -o gpurun_out/prof_tc_$TAG \
      python tools/run_keyframe_tc.py 2 > /dev/null 2>&1; echo "ncu tensor-core rc=$?"
  timeout 600 ncu --set full --clock-control none -k regex:mvacc_trace -s 1 -c 1 -f -o gpurun_out/prof_trace_$TAG \
      python tools/bench_configs.py --only-upstream > /dev/null 2>&1; echo "ncu back-trace rc=$?"
  # the variants that are not HBM bound (motion rows: V0, V1, V2 fp32 NCHW then V0, V1 bf16 NHWC, 23 launches each)
  timeout 600 ncu --set full --clock-control none -k regex:agg_nchw_tma_kernel -s 10 -c 1 -f -o gpurun_out/prof_v0_nchw_$TAG \
      python tools/bench_configs.py --only-motion > /dev/null 2>&1; echo "ncu V0 fp32 NCHW rc=$?"
  timeout 600 ncu --set full --clock-control none -k regex:agg_nhwc_tma_kernel -s 10 -c 1 -f -o gpurun_out/prof_v0_nhwc_$TAG \
      python tools/bench_configs.py --only-motion > /dev/null 2>&1; echo "ncu V0 bf16 NHWC rc=$?"
  timeout 600 ncu --set full --clock-control none -k regex:agg_nhwc_tma_kernel -s 33 -c 1 -f -o gpurun_out/prof_v1_nhwc_$TAG \
      python tools/bench_configs.py --only-motion > /dev/null 2>&1; echo "ncu V1 bf16 NHWC rc=$?"
  timeout 600 ncu --set full --clock-control none -k regex:agg_nchw_tma_kernel -s 20 -c 1 -f -o gpurun_out/prof_cfg5_$TAG \
      python tools/bench_streams.py --streams 64 > /dev/null 2>&1; echo "ncu cfg5 rc=$?"
  timeout 600 ncu --set full --clock-control none -k regex:agg_nhwc_win_kernel -s 5 -c 1 -f -o gpurun_out/prof_win_v0_$TAG \
      python tools/bench_configs.py --only-nhwc-win > /dev/null 2>&1; echo "ncu window kernel rc=$?"
  # text summaries for profiles/; gpurun brings back at most 64 MiB, so only the two main reports travel as .ncu-rep
  for r in tma:64 tc:16 trace:64 v0_nchw:64 v0_nhwc:64 v1_nhwc:64 cfg5:80 win_v0:64; do
    python tools/ncu_summary.py gpurun_out/prof_${r%%:*}_$TAG.ncu-rep gpurun_out/ncu_${r%%:*}_$TAG.txt ${r##*:} > /dev/null 2>&1
  done
  rm -f gpurun_out/prof_trace_$TAG.ncu-rep gpurun_out/prof_v0_nchw_$TAG.ncu-rep gpurun_out/prof_v0_nhwc_$TAG.ncu-rep \
        gpurun_out/prof_v1_nhwc_$TAG.ncu-rep gpurun_out/prof_cfg5_$TAG.ncu-rep gpurun_out/prof_win_v0_$TAG.ncu-rep
  ls -la gpurun_out | tail -30 ;;
sanitize)
  SEL='all_tma_kernel_every_variant or window_kernel or cooperative or host_aggregator or test_bilinear_sampler or cur_frame_path or fused_golden'
  for tool in memcheck synccheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_tc_convs.py tests/test_host_path.py -m gpu -x -q -k "$SEL or conv_bf16 or host_path_matches" > gpurun_out/sanitizer_${TAG}_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_${TAG}_$tool.log | tail -3
  done ;;
*) echo "unknown visit $WHAT"; exit 2 ;;
esac
