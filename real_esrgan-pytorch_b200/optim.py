"""Optimizer side of the training step on the flat parameter vector.

`FlatAdamEMA` replaces `optim.Adam(model.parameters(), lr, betas)` + `EMA(model, decay)` of the reference
(train_realesrnet.py:183-185, 197-200, 388-394; model.py:30-61): the generator's parameters are re-pointed to views of
ONE contiguous fp32 vector, so that forward/backward (`autograd.TrainStep`, gradients in the same flat layout), the Adam
update and the EMA update are each a single pass over 16.7 M elements (C ABI `resr_adam_ema_step`), with no per-tensor
Python loop. `state_dict` key names / shapes of the generator are unchanged (the views are ordinary parameters).
"""
import torch

from . import _lib


class FlatAdamEMA:
    def __init__(self, gen, lr: float = 2e-4, betas=(0.9, 0.99), eps: float = 1e-8, ema_decay: float = 0.999):
        params = list(gen.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.ResrError("FlatAdamEMA needs the generator on a CUDA device (no CPU fallback)")
        self.gen, self.lr, self.betas, self.eps, self.ema_decay = gen, float(lr), tuple(betas), float(eps), float(ema_decay)
        n = _lib.lib().resr_generator_num_params()
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        pos = 0
        for p in params:  # state_dict order == layout of resr_generator_tensor_span
            k = p.numel()
            self.flat[pos:pos + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[pos:pos + k].view_as(p)  # parameters become views of the flat master copy
            pos += k
        assert pos == n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.shadow = self.flat.clone()  # EMA.register(): model.py:36-39
        self._backup = None
        self.step_count = 0
        gen._flat_master = self.flat  # the module packs straight from the master copy (no gather)
        gen._packed_version = None

    def step(self, grads_flat: torch.Tensor, grad_scale: float = 1.0, lr: float = None):
        """One Adam step on `grads_flat` (the vector TrainStep returns) followed by the EMA update."""
        if grads_flat.numel() != self.flat.numel() or grads_flat.dtype != torch.float32 or not grads_flat.is_contiguous():
            raise _lib.ResrError("grads_flat must be the contiguous fp32 flat gradient vector")
        self.step_count += 1
        _lib.check(_lib.lib().resr_adam_ema_step(
            _lib.ptr(self.flat), _lib.ptr(grads_flat), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.shadow),
            self.flat.numel(), self.lr if lr is None else float(lr), self.betas[0], self.betas[1], self.eps, self.step_count,
            self.ema_decay, float(grad_scale), _lib.stream_ptr()))
        self.gen._packed_version = None  # the module repacks its tensor-core weight tiles on the next forward

    # reference EMA API (model.py:51-61)
    def apply_shadow(self):
        self._backup = self.flat.clone()
        self.flat.copy_(self.shadow)
        self.gen._packed_version = None

    def restore(self):
        if self._backup is None:
            raise _lib.ResrError("restore() without apply_shadow()")
        self.flat.copy_(self._backup)
        self._backup = None
        self.gen._packed_version = None

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "shadow": self.shadow,
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "ema_decay": self.ema_decay}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.shadow.copy_(sd["shadow"])
