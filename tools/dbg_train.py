import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, resr_b200
from oracle import generator as og
g = resr_b200.model.Generator(3, 3, 4); g.load_state_dict(og.random_state_dict(2)); g = g.cuda().train()
torch.manual_seed(11)
for shape in [(2, 16, 32), (2, 16, 24), (2, 16, 64)]:
    n, h, w = shape
    lr = torch.rand(n, 3, h, w, device="cuda"); hr = torch.rand(n, 3, 4 * h, 4 * w, device="cuda")
    loss_e, sr, flat_e = resr_b200.autograd.l1_loss_backward(g, lr, hr)
    torch.cuda.synchronize()
    x = lr.cpu(); 
    ref_loss, ref_grads, ref_sr = og.l1_loss_and_grads(x, hr.cpu(), {k: v.detach().cpu() for k, v in g.state_dict().items()})
    ref_flat = torch.cat([ref_grads[k].reshape(-1) for k in g.state_dict()])
    print(shape, "eager loss", loss_e.item(), ref_loss.item(), "norm", flat_e.norm().item(), ref_flat.norm().item(), "sr err", (sr.cpu()-ref_sr).abs().max().item())
    ts = resr_b200.autograd.TrainStep(g, n, h, w)
    for it in range(3):
        loss_g, _, flat_g = ts.step(lr, hr)
        torch.cuda.synchronize()
        print("   step", it, loss_g.item(), flat_g.norm().item(), "graph", ts.is_graph)
