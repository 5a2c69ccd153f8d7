"""Training path of the generator mirror: forward that keeps activations, fused L1 loss and the full backward pass,
all inside libresr.so (resr_generator_forward_train / _backward_l1 / _backward, include/resr.h).

Replaces, for the generator, `sr = model(lr); loss = L1(sr, hr); loss.backward()` of the reference training loop
(train_realesrnet.py:383-388). Gradients arrive as ONE flat fp32 vector in state_dict order (the layout a DDP
all-reduce or a fused optimizer wants) and are scattered into `param.grad`.
"""
import ctypes

import torch

from . import _lib


def _train_workspace(gen, n, h, w, device):
    need = _lib.lib().resr_generator_train_workspace_bytes(n, h, w)
    ws = getattr(gen, "_train_ws", None)
    if ws is None or ws.numel() < need + 1024 or ws.device != device:
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
        gen._train_ws = ws
    base = ws.data_ptr()
    off = (-base) % 1024
    return ctypes.c_void_p(base + off), ws.numel() - off


def _scatter_grads(gen, flat: torch.Tensor, accumulate: bool = True):
    pos = 0
    for p in gen.parameters():
        n = p.numel()
        g = flat[pos:pos + n].view_as(p)
        if p.requires_grad:
            if p.grad is None or not accumulate:
                p.grad = g.clone()
            else:
                p.grad.add_(g)
        pos += n


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gen, *params):
        n, _, h, w = x.shape
        if w % 8 != 0:
            raise _lib.ResrError(f"the training path needs W % 8 == 0 (got {w})")
        xc = x.detach().contiguous().float()
        gen._ensure_packed()
        y = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=x.device)
        wp, wbytes = _train_workspace(gen, n, h, w, x.device)
        _lib.check(_lib.lib().resr_generator_forward_train(gen._native(), _lib.ptr(xc), _lib.ptr(y), n, h, w, wp, wbytes,
                                                           _lib.stream_ptr()))
        ctx.gen = gen
        ctx.shape = (n, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        gen = ctx.gen
        n, h, w = ctx.shape
        dyc = dy.detach().contiguous().float()
        flat = torch.empty(_lib.lib().resr_generator_num_params(), dtype=torch.float32, device=dy.device)
        wp, wbytes = _train_workspace(gen, n, h, w, dy.device)
        _lib.check(_lib.lib().resr_generator_backward(gen._native(), _lib.ptr(dyc), _lib.ptr(flat), n, h, w, wp, wbytes,
                                                      _lib.stream_ptr()))
        grads, pos = [], 0
        for p in gen.parameters():
            grads.append(flat[pos:pos + p.numel()].view_as(p) if p.requires_grad else None)
            pos += p.numel()
        return (None, None) + tuple(grads)  # no gradient w.r.t. the LR input (the degradation output is detached)


def generator_apply(gen, x):
    return _GeneratorFn.apply(x, gen, *list(gen.parameters()))


def l1_loss_backward(gen, lr: torch.Tensor, hr: torch.Tensor, accumulate: bool = False):
    """Fused training step core: sr = G(lr); loss = mean|sr - hr|; d loss / d params. Returns (loss, sr, flat_grads);
    flat_grads (fp32, state_dict order) is also scattered into param.grad."""
    n, _, h, w = lr.shape
    if w % 8 != 0:
        raise _lib.ResrError(f"the training path needs W % 8 == 0 (got {w})")
    dev = lr.device
    xc = lr.detach().contiguous().float()
    hrc = hr.detach().contiguous().float()
    gen._ensure_packed()
    sr = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=dev)
    flat = torch.empty(_lib.lib().resr_generator_num_params(), dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    wp, wbytes = _train_workspace(gen, n, h, w, dev)
    lib = _lib.lib()
    _lib.check(lib.resr_generator_forward_train(gen._native(), _lib.ptr(xc), _lib.ptr(sr), n, h, w, wp, wbytes, _lib.stream_ptr()))
    _lib.check(lib.resr_generator_backward_l1(gen._native(), _lib.ptr(hrc), _lib.ptr(flat), _lib.ptr(loss), n, h, w, wp, wbytes,
                                              _lib.stream_ptr()))
    _scatter_grads(gen, flat, accumulate)
    return loss, sr, flat
