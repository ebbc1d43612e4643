#!/bin/bash
# One GPU visit: smoke, bench, launch list, and a full ncu capture of the dominant kernel.
# Run under gpurun from the repo root:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
kill $SMI
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nchw_tma -s 3 -c 2 -f -o gpurun_out/prof_tma_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nhwc_tma -c 2 -f -o gpurun_out/prof_nhwc_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_nhwc_$TAG.log 2>&1; echo "ncu nhwc rc=$?"
ls -la gpurun_out/
