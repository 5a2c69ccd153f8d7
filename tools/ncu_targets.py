"""Targets for the ncu captures of round 2 (development aid; numbers printed under ncu are never bench values).
    python tools/ncu_targets.py gen      -> two cfg3 forwards (capture the second: --launch-skip 353 -c 352)
    python tools/ncu_targets.py rdb      -> the same, meant for `--set full` of a few launches of one dense block
    python tools/ncu_targets.py degrade  -> three direct (non-graph) S0 batches at cfg2 (capture the last one)
    python tools/ncu_targets.py train    -> two eager (RESR_NO_GRAPH) bf16 training steps at cfg4 (capture wgrad_mn launches)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import resr_b200

what = sys.argv[1] if len(sys.argv) > 1 else "gen"
dev = torch.device("cuda")
if what == "train":
    os.environ["RESR_NO_GRAPH"] = "1"
    torch.manual_seed(0)
    g = resr_b200.model.Generator(3, 3, 4).to(dev).train()
    g.set_precision("bf16")
    lr = torch.rand(16, 3, 64, 64, device=dev)
    hr = torch.rand(16, 3, 256, 256, device=dev)
    for _ in range(2):
        loss, _, _ = resr_b200.autograd.l1_loss_backward(g, lr, hr)
    torch.cuda.synchronize()
    print("ok", float(loss))
elif what in ("gen", "rdb"):
    torch.manual_seed(0)
    g = resr_b200.model.Generator(3, 3, 4).to(dev).eval()
    x = torch.rand(64, 3, 128, 128, device=dev)
    with torch.no_grad():
        for _ in range(2):
            y = g(x)
    torch.cuda.synchronize()
    print("ok", float(y.mean()))
else:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    ip = resr_b200.imgproc
    B = 16
    plan = resr_b200.plan.canonical_plan_s0(B, 256, 256, seed=0)
    hr = torch.rand(B, 3, 256, 256, device=dev)
    k1, k2, sk = bench.s0_kernels(B, dev)
    plan_d = ip.plan_to_device(plan, dev)
    torch.cuda.synchronize()
    for it in range(3):
        torch.cuda.nvtx.range_push(f"batch{it}")
        lr, hrc = ip.degrade_batch(hr, k1, k2, sk, plan_d)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    print("ok", float(lr.mean()))
