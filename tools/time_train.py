"""Device-timed training-step core (fwd + L1 + bwd) at cfg4 shapes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resr_b200
n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (16, 64, 64)))
torch.manual_seed(0)
g = resr_b200.model.Generator(3, 3, 4).cuda().train()
g.set_precision(os.environ.get("RESR_PREC", "bf16"))
lr = torch.rand(n, 3, h, w, device="cuda"); hr = torch.rand(n, 3, 4 * h, 4 * w, device="cuda")
ts = resr_b200.autograd.TrainStep(g, n, h, w)
run = (lambda: ts.step(lr, hr, scatter=False)) if os.environ.get("RESR_GRAPH", "1") == "1" else (lambda: resr_b200.autograd.l1_loss_backward(g, lr, hr))
for _ in range(3):
    loss, sr, flat = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
t0 = time.perf_counter()
e0.record()
for _ in range(iters):
    loss, sr, flat = run()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = 3 * 35853696.0 * n * h * w
print(f"train core [{g.precision}] {n}x3x{h}x{w}: {ms:.2f} ms/step (host enqueue {1e3*(t1-t0)/iters:.2f} ms)  {n/ms*1e3:.0f} pairs/s  {flop/ms/1e9:.1f} TFLOP/s  loss {loss.item():.4f} graph={ts.is_graph}")
