"""CPU, world_size 2, gloo: the host-side logic of the data-parallel training step — the single gradient all-reduce over
the flat parameter vector and the scatter back into param.grad in state_dict order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import resr_b200
    torch.manual_seed(0)
    g = resr_b200.model.Generator(3, 3, 4)
    n = sum(p.numel() for p in g.parameters())
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)  # rank-dependent "gradients"
    resr_b200.autograd.allreduce_mean_(flat, None, world)
    resr_b200.autograd._scatter_grads(g, flat, accumulate=False)
    expect = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    got = torch.cat([p.grad.reshape(-1) for p in g.parameters()])
    ok = torch.equal(got, expect)
    last = list(g.parameters())[-1]
    ok = ok and last.grad.shape == (3,) and float(last.grad[-1]) == float(expect[-1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_scatter_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
