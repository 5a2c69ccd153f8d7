"""Per-stage device time of the canonical plan S0 at cfg2 (development aid): every stage is run `reps` times back to back on
the stage input captured from one pipeline execution (inputs L2-hot, launch gaps hidden), CUDA events around the loop."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import resr_b200

ip = resr_b200.imgproc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")
plan = ip.plan_to_device(resr_b200.plan.canonical_plan_s0(B, 256, 256, seed=0), dev)
hr = torch.rand(B, 3, 256, 256, device=dev)
k1, k2, sk = bench.s0_kernels(B, dev)
stages = []
ip.degrade_batch(hr, k1, k2, sk, plan, stages=stages)
inputs = {"usm": hr}
prev = hr
for name, t in stages:
    inputs[name] = prev
    prev = t
usm, jp = ip.USMSharp(50, 0), ip.DiffJPEG(False)
q1, q2 = plan["jpeg1_quality"], plan["jpeg2_quality"]
ops = {
    "usm": lambda x: usm(x, 0.5, 10),
    "blur1": lambda x: ip.filter2d_torch(x, k1),
    "resize1": lambda x: ip.interpolate(x, scale_factor=0.5, mode="bicubic"),
    "noise1": lambda x: ip._noise(x, plan["noise1"]),
    "jpeg1": lambda x: jp(x, q1.clone(), clamp_input=True),
    "blur2": lambda x: ip.filter2d_torch(x, k2),
    "resize2": lambda x: ip.interpolate(x, size=(64, 64), mode="bilinear"),
    "noise2": lambda x: ip._noise(x, plan["noise2"]),
    "sinc": lambda x: ip.filter2d_torch(x, sk),
    "jpeg2": lambda x: jp(x, q2.clone(), clamp_input=True),
    "crop": lambda x: ip._crop(x, 0, 0, 64, 64, round_to_u8=True),
}
reps = 30
total = 0.0
for name, fn in ops.items():
    x = inputs[name if name != "crop" else "jpeg2"] if name != "crop" else stages[-1][1]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(x)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn(x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * reps) * 1e3
    total += us
    print(f"{name:8s} {tuple(x.shape)!s:20s} {us:7.1f} us")
print(f"sum {total:.1f} us")
