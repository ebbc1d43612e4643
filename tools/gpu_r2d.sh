#!/bin/bash
# Round-2 GPU visit D: profiler evidence.  Launch list of the bench command, ncu --set full of the tensor-core kernels,
# the cooperative single-frame kernel and the coviar back-trace.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 4 -c 4 -f -o gpurun_out/r2_prof_tc \
    python tools/run_keyframe_tc.py 2 > gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nchw_tma -c 1 -f -o gpurun_out/r2_prof_coop \
    python tools/bench_configs.py --only-single > gpurun_out/r2_ncu_coop.log 2>&1; echo "ncu coop rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mvacc_trace -s 1 -c 1 -f -o gpurun_out/r2_prof_trace \
    python tools/bench_configs.py --only-upstream > gpurun_out/r2_ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nchw_tma -s 3 -c 1 -f -o gpurun_out/r2_prof_tma \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/r2_ncu_tma.log 2>&1; echo "ncu tma rc=$?"
ls -la gpurun_out/*.ncu-rep
