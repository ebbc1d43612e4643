#!/bin/bash
# Rebuild with each stream cache policy on the GPU box and bench the headline kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pol in 0 1 2 3; do
  LSFA_NVCC_EXTRA="-DLSFA_LD_POLICY=$pol" python -m lsfa_b200._build --force > /dev/null 2>&1
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra --e2e-steps 2 > gpurun_out/bench_pol$pol.json 2>gpurun_out/bench_pol$pol.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pol$pol.json"))
print("policy $pol: ms_per_step %.4f  GB/s %.0f  frac %.3f" % (d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"]))
PY
done
LSFA_NVCC_EXTRA="" python -m lsfa_b200._build --force > /dev/null 2>&1
