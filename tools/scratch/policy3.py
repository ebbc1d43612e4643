import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lsfa_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench_configs import synth, time_ms, peak
dev = torch.device("cuda", 0)
N, C, H, W = 64, 1024, 38, 63
d = synth(N, C, H, W, 600, 1000, dev, max_px=96)
flow = ops.mv_pool(d["mv"])
og = torch.randn_like(d["key"]); gk = torch.empty_like(d["key"]); gf = torch.empty_like(flow)
ws = torch.empty(ops.A.load().lsfa_bilinear_sampler_backward_workspace_bytes(N, C, H, W, H, W), dtype=torch.uint8, device=dev)
for pct in (0, 5, 10, 25, 100):
    os.environ["LSFA_TMA_POOL_PCT"] = str(pct)
    a = time_ms(lambda: ops.warp_backward(d["key"], flow, og, grad_key=gk, grad_flow=gf, workspace=ws, kernel="gather"), 3, 20)
    b = time_ms(lambda: ops.warp_backward(d["key"], flow, og, grad_key=gk, req_flow="null", workspace=ws, kernel="gather"), 3, 20)
    print("pool %3d%%: both %.4f ms, key only %.4f ms" % (pct, a, b), flush=True)
