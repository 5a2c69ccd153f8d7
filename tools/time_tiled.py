"""cfg5 (BASELINE.json configs[4]): 1x3x2048x2048 -> 1x3x8192x8192 by halo tiles on one GPU (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resr_b200
torch.set_grad_enabled(False)
torch.manual_seed(0)
g = resr_b200.model.Generator(3, 3, 4).cuda().eval()
x = torch.rand(1, 3, 2048, 2048, device="cuda")
out = torch.empty(1, 3, 8192, 8192, device="cuda")
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y, tiles = resr_b200.model.infer_tiled(g, x, tile_h=512, tile_w=1024, halo=16, out=out)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"pass {it}: {len(tiles)} tiles, {dt*1e3:.1f} ms, {2048*2048/dt/1e6:.2f} LR Mpix/s, finite={bool(torch.isfinite(y).all())}, "
          f"range [{y.min().item():.3f}, {y.max().item():.3f}], mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
