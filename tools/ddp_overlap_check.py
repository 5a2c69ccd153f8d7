"""torchrun check (N >= 2 GPUs): the bucketed, overlapped gradient all-reduce of TrainStep gives the same flat gradient vector as
the single all-reduce after the step, in the eager step and in the CUDA-graph replay; prints the device time of both.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_overlap_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import resr_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("RESR_NCCL_HIGH_PRIO"):
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=dev, pg_options=opts)
else:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
gen = resr_b200.model.Generator(3, 3, 4).to(dev).train()
gen.set_precision(os.environ.get("RESR_PREC", "bf16"))
n, h, w = 16, 64, 64
g = torch.Generator().manual_seed(100 + rank)
lr = torch.rand(n, 3, h, w, generator=g).to(dev)
hr = torch.rand(n, 3, 4 * h, 4 * w, generator=g).to(dev)
ts = resr_b200.autograd.TrainStep(gen, n, h, w, dev, None, world)
res = {}
for overlap in (True, False, True):
    ts.overlap = overlap
    for _ in range(3):
        loss, _, flat = ts.step(lr, hr, scatter=False)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        loss, _, flat = ts.step(lr, hr, scatter=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[overlap] = (flat.clone(), ms.item())
    if rank == 0:
        print(f"overlap={overlap}: {ms.item():.3f} ms per step (max over {world} ranks), graph={ts.is_graph}, loss {loss.item():.5f}", flush=True)
d = (res[True][0] - res[False][0]).abs().max().item()
ref = res[False][0].abs().max().item()
# every rank must hold the same reduced vector
chk = res[True][0].double().sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(abs(c.item() - allc[0].item()) == 0 for c in allc)
if rank == 0:
    print(f"max |overlapped - single all-reduce| = {d:.3e} (gradient scale {ref:.3e}); identical across ranks: {same}", flush=True)
    print("DDP_OVERLAP_OK" if d == 0.0 and same else "DDP_OVERLAP_MISMATCH", flush=True)
dist.destroy_process_group()
