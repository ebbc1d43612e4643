import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lsfa_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench_configs import synth, time_ms, peak
dev = torch.device("cuda", 0)
pk = peak()
s = torch.cuda.current_stream().cuda_stream
N, C, H, W = 64, 1024, 38, 63
HW, F4 = H * W, C * H * W * 4
d = synth(N, C, H, W, 600, 1000, dev)
flow = ops.mv_pool(d["mv"])
cases = {
 "V0": (dict(), 2 * F4),
 "Scale": (dict(scale_map=d["scale_map"]), 3 * F4),
 "V1": (dict(cur=d["cur"], res=d["res"], rnet_w=d["rnet_w"], rnet_b=d["rnet_b"], weight_mode="add"), 3 * F4),
 "V2": (dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode="logits", logits=d["logits"]), 4 * F4),
}
for fk, fl in (("flow", flow), ("raw", d["mv"])):
    for name, (kw, b) in cases.items():
        out = []
        for st_, nr in ((0, 0), (1, 0), (0, 1), (1, 1)):
            os.environ.pop("LSFA_TMA_STATIC", None); os.environ.pop("LSFA_TMA_NO_RECORDS", None)
            if st_: os.environ["LSFA_TMA_STATIC"] = "1"
            if nr: os.environ["LSFA_TMA_NO_RECORDS"] = "1"
            p = ops.PreparedAggregate(d["key"], fl, flow_kind=fk, **kw)
            ms = time_ms(lambda: p.run(s), 5, 40)
            out.append("%s%s %.4f (%.3f)" % ("static" if st_ else "dyn", "+inline" if nr else "+records", ms, N * b / (ms / 1e3) / 1e9 / pk))
        print(fk, name, " | ".join(out), flush=True)
