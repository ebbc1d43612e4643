import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lsfa_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench_configs import synth, time_ms, peak
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
if which == "cfg4":
    N, C, H, W, mvh, mvw, mp = 64, 1024, 68, 120, 1080, 1920, 96
else:
    N, C, H, W, mvh, mvw, mp = 128, 1024, 38, 63, 600, 1000, 32
d = synth(N, C, H, W, mvh, mvw, dev, max_px=mp)
nh = {k: ops.to_nhwc(d[k], torch.bfloat16) for k in ("key", "cur", "scale_map")}
p = ops.PreparedAggregate(nh["key"], d["mv"], flow_kind="raw", cur=nh["cur"], scale_map=nh["scale_map"],
                          weight_mode="logits", logits=d["logits"], layout="nhwc_bf16")
ms = time_ms(lambda: p.run(s), 3, 10)
b = N * (4 * C * H * W * 2 + 40 * H * W)
print(which, "ms", ms, "GB/s", b / ms / 1e6, "frac", b / ms / 1e6 / peak())
