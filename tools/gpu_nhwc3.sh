#!/bin/bash
# producer-warp count experiment (libraries built with -DLSFA_NT_PRODUCERS=n) + ncu of the residual variant
TAG=${1:-n}
mkdir -p gpurun_out
show() { python - "$1" <<PY
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("    %-62s %7.4f ms %7.0f GB/s  %.3f" % (d["config"][:62], d["ms_per_step"], d["achieved_gbs"], d["frac_of_measured_peak"]))
PY
}
timeout 600 python -m pytest tests -m gpu -x -q -k "nhwc or cur_frame_path or identical_bits" 2>&1 | tail -3
echo "== NP=2 (default)"; timeout 600 python tools/bench_configs.py --only-nhwc > gpurun_out/nhwc_$TAG.jsonl 2> gpurun_out/nhwc_$TAG.err; show gpurun_out/nhwc_$TAG.jsonl
for np in 3 4; do
  echo "== NP=$np"; LSFA_B200_LIB=$PWD/lsfa_b200/lib/liblsfa_np$np.so timeout 600 python -m pytest tests -m gpu -x -q -k "nhwc_all_tma" 2>&1 | tail -1
  LSFA_B200_LIB=$PWD/lsfa_b200/lib/liblsfa_np$np.so timeout 600 python tools/bench_configs.py --only-nhwc-tma 2>/dev/null | sort -u > gpurun_out/nhwc_${TAG}_np$np.jsonl; show gpurun_out/nhwc_${TAG}_np$np.jsonl
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg_nhwc_tma -s 41 -c 1 -f -o gpurun_out/prof_nhwc_tma_res_$TAG \
    python tools/bench_configs.py --only-nhwc-tma --quick > gpurun_out/ncu_nhwc_tma_res_$TAG.log 2>&1; echo "ncu rc=$?"
