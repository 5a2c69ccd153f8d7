"""'Library bar' (SURVEY.md §8d): the reference architecture run through stock PyTorch / cuDNN on the same GPU.
Development aid only: uses the oracle's torch restatement of the network, never part of the product path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import generator as og
torch.manual_seed(0)
sd = og.random_state_dict(0)
n, h, w = 64, 128, 128
x = torch.rand(n, 3, h, w, device="cuda")
sd = {k: v.cuda() for k, v in sd.items()}
def run(dtype, cl, bs):
    xs = x[:bs]
    if cl:
        xs = xs.contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
        return og.generator_forward(xs, sd)
for name, dtype, cl in [("fp32 (TF32 off)", None, False), ("bf16 autocast channels_last", torch.bfloat16, True), ("fp16 autocast channels_last", torch.float16, True)]:
    torch.backends.cudnn.benchmark = True
    for bs in (16,):
        try:
            for _ in range(2):
                run(dtype, cl, bs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                run(dtype, cl, bs)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"library bar {name}: batch {bs}x3x{h}x{w}  {ms:.1f} ms  {bs*h*w/ms/1e3:.2f} LR Mpix/s")
        except Exception as e:
            print(f"library bar {name}: failed {type(e).__name__}: {str(e)[:100]}")

# cfg4: the training-step core (forward + L1 + backward) of the same network through stock PyTorch autograd / cuDNN
n4, h4, w4 = 16, 64, 64
x4 = torch.rand(n4, 3, h4, w4, device="cuda")
hr4 = torch.rand(n4, 3, 4 * h4, 4 * w4, device="cuda")
for name, dtype, cl in [("fp32 (TF32 off)", None, False), ("bf16 autocast channels_last", torch.bfloat16, True)]:
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    xs = x4.contiguous(memory_format=torch.channels_last) if cl else x4

    def step():
        for p in params.values():
            p.grad = None
        with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
            sr = og._forward(xs, params, 23)
            loss = torch.nn.functional.l1_loss(sr.float(), hr4)
        loss.backward()
        return loss
    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"library bar cfg4 training step {name}: {n4}x3x{h4}x{w4}  {ms:.1f} ms  {n4/ms*1e3:.0f} pairs/s")
    except Exception as e:
        print(f"library bar cfg4 {name}: failed {type(e).__name__}: {str(e)[:100]}")
