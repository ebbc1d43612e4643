#!/bin/bash
# Round-2 GPU visit B: host path tests, bench line with the new rows, regenerate the cuDNN fixture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_path.py tests/test_cudnn_pin.py -q -m gpu --tb=short 2>&1 | tail -30
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "host_aggregator" --tb=short 2>&1 | tail -10
timeout 120 python tools/make_golden_cudnn.py 2>&1 | tail -2
( time python bench.py --steps 50 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2>&1 | tail -4; echo "bench rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2b_bench.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')})
print('roofline', {k:l['roofline'][k] for k in ('achieved','frac','traffic','traffic_source')})
print('e2e', json.dumps(l['e2e'])[:1500])
print('cpu', l['cpu_baseline'])
for k,v in l.get('extra',{}).items(): print(k, json.dumps(v)[:700])
"; tail -5 gpurun_out/r2b_bench.err
( time python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2b_bench_ref.json ) 2>&1 | tail -4; cut -c1-400 gpurun_out/r2b_bench_ref.json
