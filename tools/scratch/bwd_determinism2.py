import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from lsfa_b200 import ops
from lsfa_b200 import _cabi as A
dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
N, C, H, W = 8, 64, 38, 63
HW = H * W
data = torch.from_numpy(rng.standard_normal((N, C, H, W), dtype=np.float32)).to(dev)
og = torch.from_numpy(rng.standard_normal((N, C, H, W), dtype=np.float32)).to(dev)
flow = torch.zeros((N, 2, H, W), device=dev)
need = A.load().lsfa_bilinear_sampler_backward_workspace_bytes(N, C, H, W, H, W)
r16 = lambda v: (v + 15) // 16 * 16
o_rec = r16(N * 4) + 16
o_off = o_rec + r16(N * HW * 16)
o_w = o_off + r16(N * 8 * HW * 2)
o_cnt = o_w + r16(N * 8 * HW * 4)
o_ovf = o_cnt + r16(N * HW)
outs, wss = [], []
for it in range(6):
    ws = torch.zeros(need, dtype=torch.uint8, device=dev)
    gk, gf = ops.warp_backward(data, flow, og, workspace=ws, kernel="gather")
    torch.cuda.synchronize()
    outs.append(gk.clone()); wss.append(ws.clone())
L = 8
for it in range(1, 6):
    ne = (outs[it].view(torch.int32) != outs[0].view(torch.int32))
    print("run", it, "out diffs", int(ne.sum()), "ovf_count", int(wss[it][r16(N*4):r16(N*4)+4].view(torch.int32)[0]),
          "cnt diff", int((wss[it][o_cnt:o_ovf] != wss[0][o_cnt:o_ovf]).sum()),
          "off diff", int((wss[it][o_off:o_w] != wss[0][o_off:o_w]).sum()),
          "w diff", int((wss[it][o_w:o_cnt] != wss[0][o_w:o_cnt]).sum()),
          "rec diff", int((wss[it][o_rec:o_off] != wss[0][o_rec:o_off]).sum()))
    if ne.any():
        n, c, y, x = ne.nonzero()[0].tolist()
        q = y * W + x
        cnt = wss[0][o_cnt:o_ovf].view(N, HW)[n, q].item()
        offs = wss[0][o_off:o_w].view(torch.int16).view(N, L, HW)[n, :, q].tolist()
        wts = wss[0][o_w:o_cnt].view(torch.float32).view(N, L, HW)[n, :, q].tolist()
        print("  first diff at", (n, c, y, x), "cnt", cnt, "offs", [(o & 0xffff) >> 2 for o in offs], "taps", [o & 3 for o in offs], "w", wts)
        print("  values", outs[0][n, c, y, x].item(), outs[it][n, c, y, x].item(), "og", og[n, c, y, x].item())
