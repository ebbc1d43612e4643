#!/bin/bash
# Round-2 GPU visit C: cooperative one-launch form, single-tile tail of the tensor-core kernel, direct-store experiment.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -12 ) 2>&1 | tail -16
timeout 300 python tools/bench_configs.py --only-single > gpurun_out/r2c_single.jsonl 2>gpurun_out/r2c_single.err; cut -c1-200 gpurun_out/r2c_single.jsonl; tail -3 gpurun_out/r2c_single.err
LSFA_TMA_NO_COOP=1 timeout 300 python tools/bench_configs.py --only-single 2>/dev/null | cut -c1-160 | head -4
timeout 400 python tools/bench_configs.py --only-keyframe 2>/dev/null | grep tcgen05 | cut -c1-230
echo "--- cfg5 staged vs direct-with-cur"
timeout 300 python tools/bench_streams.py --streams 256 | cut -c1-400
LSFA_TMA_DIRECT_CUR=1 timeout 300 python tools/bench_streams.py --streams 256 | cut -c1-400
echo "--- headline staged vs direct-with-cur"
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra --e2e-steps 2 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(l['value'], l['roofline']['frac'])"
LSFA_TMA_DIRECT_CUR=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra --e2e-steps 2 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(l['value'], l['roofline']['frac'])"
