import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lsfa_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench_configs import synth
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream().cuda_stream
for (N, C, H, W, mvh, mvw) in ((3, 8, 68, 120, 1080, 1920), (2, 4, 100, 132, 1600, 2112), (16, 256, 68, 120, 1080, 1920), (64, 1024, 38, 63, 600, 1000), (5, 64, 38, 63, 600, 1000)):
    d = synth(N, C, H, W, mvh, mvw, dev, max_px=96)
    byp = torch.zeros(N, dtype=torch.uint8, device=dev); byp[::3] = 1
    for name, kw in (("V2", dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode="logits", logits=d["logits"])),
                     ("V2+byp", dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode="logits", logits=d["logits"], bypass=byp)),
                     ("V1", dict(cur=d["cur"], res=d["res"], rnet_w=d["rnet_w"], rnet_b=d["rnet_b"], weight_mode="add")),
                     ("V0", dict())):
        os.environ["LSFA_TMA_POOL_PCT"] = "0"
        ref = ops.warp_scale_aggregate(d["key"], d["mv"], flow_kind="raw", **kw).clone()
        bad = 0
        for pct in (10, 50, 100):
            os.environ["LSFA_TMA_POOL_PCT"] = str(pct)
            for it in range(40):
                out = ops.warp_scale_aggregate(d["key"], d["mv"], flow_kind="raw", **kw)
                if not torch.equal(out, ref):
                    bad += 1
        print((N, C, H, W), name, "mismatching runs:", bad, "of 120", flush=True)
