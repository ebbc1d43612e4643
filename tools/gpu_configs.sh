#!/bin/bash
# parity tests + the per-config kernel table.  Usage: bash tools/gpu_configs.sh tag [quick]
TAG=${1:-c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/bench_configs.py ${2:+--quick} > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "rc=$?"
python - <<PY
import json
for l in open("gpurun_out/configs_$TAG.jsonl"):
    d=json.loads(l); print("%-62s %7.4f ms %10.0f fps %7.0f GB/s  %.3f" % (d["config"][:62], d["ms_per_step"], d["frames_per_s"], d["achieved_gbs"], d["frac_of_measured_peak"]))
PY
tail -3 gpurun_out/configs_$TAG.err
