"""Stress of the channels-last all-TMA kernel: many back-to-back launches on full-size inputs (claimed and static batch
orders, bf16 and fp32, blend and residual variants); every output must equal the tile kernel's bit for bit."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lsfa_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench_configs import synth

dev = torch.device("cuda", 0)
bad = 0
total = 0
for (N, H, W, mvh, mvw) in ((48, 38, 63, 600, 1000), (12, 68, 120, 1080, 1920), (7, 17, 23, 272, 368)):
    d = synth(N, 1024, H, W, mvh, mvw, dev, max_px=96)
    for dt, lay in ((torch.bfloat16, "nhwc_bf16"), (torch.float32, "nhwc_f32")):
        nh = {k: ops.to_nhwc(d[k], dt) for k in ("key", "cur", "scale_map")}
        for name, kw in (("blend", dict(cur=nh["cur"], scale_map=nh["scale_map"], weight_mode="logits", logits=d["logits"])),
                         ("res", dict(cur=nh["cur"], res=d["res"], rnet_w=d["rnet_w"], rnet_b=d["rnet_b"], weight_mode="add")),
                         ("warp", dict())):
            ref = ops.warp_scale_aggregate(nh["key"], d["mv"], flow_kind="raw", layout=lay, force_generic=1, **kw)
            for ws in (None, False):
                p = ops.PreparedAggregate(nh["key"], d["mv"], flow_kind="raw", layout=lay, force_generic=3, workspace=ws, **kw)
                for it in range(40):
                    out = p.run()
                    total += 1
                    if not torch.equal(out.view(torch.int16 if dt == torch.bfloat16 else torch.int32),
                                       ref.view(torch.int16 if dt == torch.bfloat16 else torch.int32)):
                        bad += 1
                        print("MISMATCH", N, H, W, lay, name, ws, it, flush=True)
        del nh
    del d
    torch.cuda.empty_cache()
torch.cuda.synchronize()
print("stress: %d launches, %d mismatches" % (total, bad))
sys.exit(1 if bad else 0)
