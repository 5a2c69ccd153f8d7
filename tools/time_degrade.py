"""Device-timed degradation pipeline on cfg2 (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import resr_b200
ip = resr_b200.imgproc
B, H, W = 16, 256, 256
dev = torch.device("cuda")
plan = resr_b200.plan.canonical_plan_s0(B, H, W, seed=0)
hr = torch.rand(B, 3, H, W, device=dev)
k = torch.zeros(B, 21, 21); ax = torch.arange(21) - 10.0
for i in range(B):
    ks = 7 + 2 * (i % 8); s = 0.5 + 0.3 * i
    kk = torch.exp(-(ax[:, None] ** 2 + ax[None] ** 2) / (2 * s * s))
    kk[(ax.abs() > ks // 2)[:, None] | (ax.abs() > ks // 2)[None]] = 0
    k[i] = kk / kk.sum()
k1 = k.to(dev); k2 = k.flip(0).contiguous().to(dev)
sk = torch.zeros(B, 21, 21); sk[:, 10, 10] = 1; sk = sk.to(dev)
plan_d = ip.plan_to_device(plan, dev)
for name, fn in (("direct", None), ("graph", True)):
    if fn is None:
        run = lambda: ip.degrade_batch(hr, k1, k2, sk, plan_d)
    else:
        pipe = ip.DegradePipeline(hr, k1, k2, sk, plan)
        run = pipe
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    print(f"degrade_batch 16x3x256x256 S0 [{name}]: {e0.elapsed_time(e1)/50*1e3:.1f} us/batch")
