import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch, torch.nn.functional as F
from test_conv_gpu import _run_conv, _nhwc16
torch.manual_seed(7)
dev = "cuda"
n, h, w, cin, cout = 2, 10, 128, 192, 64
x = torch.randn(n, cin, h, w, device=dev).bfloat16().float()
wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.03).bfloat16().float()
b = torch.randn(cout, device=dev) * 0.1
x16 = _nhwc16(x, cin, torch.bfloat16)
conv = F.conv2d(x, wt, b, padding=1).permute(0, 2, 3, 1).contiguous()
r1 = torch.randn(n, h, w, cout, device=dev)
r2 = torch.randn(n, h, w, cout, device=dev)
for it in range(3):
    outf = torch.zeros(n, h, w, cout, device=dev)
    _run_conv(x16, cin, wt, b, fmt=1, ep_mode=2, res1=r1, res2=r2, outf=outf)
    torch.cuda.synchronize()
    ref = (conv * 0.2 + r1) * 0.2 + r2
    bad = (outf - ref).abs() > 2e-4
    print("bad count", bad.sum().item())
    idx = bad.nonzero()
    if len(idx):
        print("n", idx[:, 0].unique().tolist(), "y", idx[:, 1].unique().tolist())
        print("x", idx[:, 2].unique().tolist()[:40], "c", idx[:, 3].unique().tolist()[:70])
        i0 = idx[0].tolist()
        print(i0, outf[tuple(i0)].item(), ref[tuple(i0)].item(), (conv*0.2*0.2 + r2)[tuple(i0)].item(), r1[tuple(i0)].item())
        # which r1 row would explain it?
        nn_, yy, xx, cc = i0
        want = (outf[nn_, yy, xx, cc] - r2[nn_, yy, xx, cc] - conv[nn_, yy, xx, cc] * 0.04) / 0.2
        m = (r1 - want).abs() < 1e-3
        print("explained by r1 at", m.nonzero()[:5].tolist())
