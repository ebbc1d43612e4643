#!/bin/bash
# channels-last all-TMA kernel: parity tests, then the NHWC rows.  Usage: bash tools/gpu_nhwc.sh tag [full]
TAG=${1:-n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "nhwc or res_coviar or bf16 or cur_frame_path or identical_bits" 2>&1 | tail -15
if [ "$2" == "full" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
timeout 600 python tools/bench_configs.py --only-nhwc > gpurun_out/nhwc_$TAG.jsonl 2> gpurun_out/nhwc_$TAG.err; echo "rc=$?"
python - <<PY
import json
for l in open("gpurun_out/nhwc_$TAG.jsonl"):
    d=json.loads(l); print("%-66s %7.4f ms %10.0f fps %7.0f GB/s  %.3f" % (d["config"][:66], d["ms_per_step"], d["frames_per_s"], d["achieved_gbs"], d["frac_of_measured_peak"]))
PY
tail -3 gpurun_out/nhwc_$TAG.err
