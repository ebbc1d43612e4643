#!/bin/bash
# channels-last all-TMA kernel experiments: knobs + one ncu capture.  Usage: bash tools/gpu_nhwc2.sh tag
TAG=${1:-n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "nhwc or cur_frame_path or identical_bits" 2>&1 | tail -4
show() { python - "$1" <<PY
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("    %-62s %7.4f ms %7.0f GB/s  %.3f" % (d["config"][:62], d["ms_per_step"], d["achieved_gbs"], d["frac_of_measured_peak"]))
PY
}
echo "== default"; timeout 600 python tools/bench_configs.py --only-nhwc > gpurun_out/nhwc_$TAG.jsonl 2> gpurun_out/nhwc_$TAG.err; show gpurun_out/nhwc_$TAG.jsonl
for knob in ${KNOBS:-"LSFA_NT_G=2" "LSFA_NT_STAGES=2"}; do
  echo "== $knob"; env $knob timeout 600 python tools/bench_configs.py --only-nhwc-tma --quick 2>/dev/null | sort -u > gpurun_out/nhwc_${TAG}_k.jsonl; show gpurun_out/nhwc_${TAG}_k.jsonl
done
echo "== default quick"; timeout 600 python tools/bench_configs.py --only-nhwc-tma --quick 2>/dev/null | sort -u > gpurun_out/nhwc_${TAG}_k.jsonl; show gpurun_out/nhwc_${TAG}_k.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nhwc_tma -c 1 -f -o gpurun_out/prof_nhwc_tma_$TAG \
    python tools/bench_configs.py --only-nhwc-tma --quick > gpurun_out/ncu_nhwc_tma_$TAG.log 2>&1; echo "ncu rc=$?"
echo "== NCHW variants without cur: staged TMA store (default)"; timeout 300 python tools/bench_configs.py --only-nocur 2>/dev/null > gpurun_out/nocur_${TAG}_a.jsonl; show gpurun_out/nocur_${TAG}_a.jsonl
echo "== NCHW variants without cur: LSFA_TMA_DIRECT_STORE=1"; LSFA_TMA_DIRECT_STORE=1 timeout 300 python tools/bench_configs.py --only-nocur 2>/dev/null > gpurun_out/nocur_${TAG}_b.jsonl; show gpurun_out/nocur_${TAG}_b.jsonl
LSFA_TMA_DIRECT_STORE=1 timeout 600 python -m pytest tests -m gpu -x -q -k "all_tma_kernel_every_variant or bilinear_sampler or identical_bits or row_trimmed" 2>&1 | tail -2
echo "== full suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
