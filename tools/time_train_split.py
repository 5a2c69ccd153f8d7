"""Experiment: is the cfg4 training step bound by the latency of its ~1,050 dependent launches? Two independent half-batch
steps (two Generator handles with the same weights, grids sized for half the SMs via RESR_NUM_SMS) on two streams vs one
full-batch step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resr_b200

parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, h, w = 16 // parts, 64, 64
torch.manual_seed(0)
gens, steps, streams = [], [], []
for i in range(parts):
    g = resr_b200.model.Generator(3, 3, 4).cuda().train()
    g.set_precision("bf16")
    gens.append(g)
    steps.append(resr_b200.autograd.TrainStep(g, n, h, w))
    streams.append(torch.cuda.Stream())
lr = [torch.rand(n, 3, h, w, device="cuda") for _ in range(parts)]
hr = [torch.rand(n, 3, 4 * h, 4 * w, device="cuda") for _ in range(parts)]
cur = torch.cuda.current_stream()


def run():
    for i in range(parts):
        streams[i].wait_stream(cur)
        with torch.cuda.stream(streams[i]):
            steps[i].step(lr[i], hr[i], scatter=False)
    for s in streams:
        cur.wait_stream(s)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{parts} concurrent step(s) of {n}x3x{h}x{w} (RESR_NUM_SMS={os.environ.get('RESR_NUM_SMS', 'all')}): {e0.elapsed_time(e1) / 5:.2f} ms per 16 pairs")
