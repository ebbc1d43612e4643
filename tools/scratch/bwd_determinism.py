import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from lsfa_b200 import ops
dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
for (N, C) in ((4, 16), (8, 64), (64, 256)):
    data = torch.from_numpy(rng.standard_normal((N, C, 38, 63), dtype=np.float32)).to(dev)
    og = torch.from_numpy(rng.standard_normal((N, C, 38, 63), dtype=np.float32)).to(dev)
    flow = torch.zeros((N, 2, 38, 63), device=dev)
    for n in range(N):
        flow[n, 0], flow[n, 1] = [0.37, -1.6, 2.25, 0.0][n % 4], [-0.21, 0.4, 3.5, 0.0][n % 4]
    for req_flow in ("write", "null"):
        ref = None
        bad = 0
        worst = 0.0
        for it in range(30):
            gk, gf = ops.warp_backward(data, flow, og, req_flow=req_flow, kernel="gather")
            torch.cuda.synchronize()
            if ref is None:
                ref = gk.clone()
            else:
                ne = (gk.view(torch.int32) != ref.view(torch.int32))
                if ne.any():
                    bad += 1
                    worst = max(worst, float((gk - ref).abs().max()))
                    idx = ne.nonzero()[:3].tolist()
        print("N=%d C=%d req_flow=%s: %d/29 runs differ, worst abs diff %g %s" % (N, C, req_flow, bad, worst, idx if bad else ""))
