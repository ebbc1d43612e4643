import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lsfa_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench_configs import synth, time_ms, peak
dev = torch.device("cuda", 0)
pk = peak()
s = torch.cuda.current_stream().cuda_stream
def run(N, C, H, W, mvh, mvw, max_px, label):
    HW, F4 = H * W, C * H * W * 4
    d = synth(N, C, H, W, mvh, mvw, dev, max_px=max_px)
    byp = torch.zeros(N, dtype=torch.uint8, device=dev); byp[::5] = 1
    cases = {
     "V0": (dict(), 2 * F4),
     "Scale": (dict(scale_map=d["scale_map"]), 3 * F4),
     "V1": (dict(cur=d["cur"], res=d["res"], rnet_w=d["rnet_w"], rnet_b=d["rnet_b"], weight_mode="add"), 3 * F4),
     "V2": (dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode="logits", logits=d["logits"]), 4 * F4),
     "V2+bypass": (dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode="logits", logits=d["logits"], bypass=byp), 4 * F4),
    }
    for name, (kw, b) in cases.items():
        out = []
        for pct in (0, 10, 25, 50, 100):
            os.environ["LSFA_TMA_POOL_PCT"] = str(pct)
            p = ops.PreparedAggregate(d["key"], d["mv"], flow_kind="raw", **kw)
            ms = time_ms(lambda: p.run(s), 5, 40)
            out.append("pool%3d%% %.4f (%.3f)" % (pct, ms, N * b / (ms / 1e3) / 1e9 / pk))
        print(label, name, " | ".join(out), flush=True)
run(64, 1024, 38, 63, 600, 1000, 32, "38x63")
run(64, 1024, 68, 120, 1080, 1920, 96, "68x120")
