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
