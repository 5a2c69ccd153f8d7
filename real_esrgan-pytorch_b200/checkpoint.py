"""Checkpoint round trip in the reference's own file format (train_realesrnet.py:66-88 load, :117-129 save):

    {"epoch", "best_niqe", "state_dict", "ema_state_dict", "optimizer", "scheduler"}   -> g_epoch_N.pth.tar

* `state_dict` is the generator's (same key set as the reference), so `inference.py:32-33` loads our files
  and we load the reference's.
* `ema_state_dict`: the reference saves `EMA(model).state_dict()`, i.e. the wrapped model's parameters under a
  "model." prefix (its shadow dict is a plain attribute and is NOT in the file). We write the same keys and, in addition,
  the EMA shadow under "ema_shadow" so that training resumes without losing the average.
* `optimizer`: `torch.optim.Adam.state_dict()` layout (state[i] = {step, exp_avg, exp_avg_sq}, param_groups) built from /
  scattered into the flat vectors of `optim.FlatAdamEMA`, so either side can resume the other's run.
"""
import torch

from . import _lib


def _spans(gen):
    pos = 0
    for name, p in gen.named_parameters():
        yield name, pos, p.numel(), tuple(p.shape)
        pos += p.numel()


def optimizer_state_dict(gen, opt) -> dict:
    """FlatAdamEMA -> torch.optim.Adam.state_dict() of an Adam built over `gen.parameters()`."""
    state = {}
    for i, (_, pos, n, shape) in enumerate(_spans(gen)):
        state[i] = {"step": torch.tensor(float(opt.step_count)),
                    "exp_avg": opt.exp_avg[pos:pos + n].view(shape).clone(),
                    "exp_avg_sq": opt.exp_avg_sq[pos:pos + n].view(shape).clone()}
    group = {"lr": opt.lr, "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
             "params": list(range(len(state)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(gen, opt, sd: dict):
    """torch.optim.Adam.state_dict() (reference run, or ours) -> the flat vectors of FlatAdamEMA."""
    st = sd["state"]
    step = 0
    for i, (_, pos, n, _) in enumerate(_spans(gen)):
        if i not in st:
            continue  # a parameter that never received a gradient
        opt.exp_avg[pos:pos + n].copy_(st[i]["exp_avg"].reshape(-1))
        opt.exp_avg_sq[pos:pos + n].copy_(st[i]["exp_avg_sq"].reshape(-1))
        step = max(step, int(float(st[i]["step"])))
    opt.step_count = step
    if sd.get("param_groups"):
        g0 = sd["param_groups"][0]
        opt.lr, opt.betas, opt.eps = float(g0["lr"]), tuple(g0["betas"]), float(g0["eps"])


def save_checkpoint(path: str, gen, opt=None, epoch: int = 0, best_niqe: float = 100.0, scheduler_state: dict = None):
    sd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    ckpt = {"epoch": int(epoch), "best_niqe": float(best_niqe), "state_dict": sd,
            "ema_state_dict": {"model." + k: v for k, v in sd.items()},
            "optimizer": optimizer_state_dict(gen, opt) if opt is not None else None,
            "scheduler": scheduler_state}
    if opt is not None:
        ckpt["ema_shadow"] = {name: opt.shadow[pos:pos + n].view(shape).clone() for name, pos, n, shape in _spans(gen)}
    torch.save(ckpt, path)
    return ckpt


def load_checkpoint(path_or_dict, gen, opt=None, map_location=None) -> dict:
    """Loads a reference-format checkpoint (ours or the reference's) into `gen` (and `opt`). Keys may carry the
    "model." prefix of the EMA wrapper (inference.py:33 strips it the same way). Returns the checkpoint dict."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location=map_location or (lambda s, l: s))
    gen.load_state_dict({k.replace("model.", ""): v for k, v in ckpt["state_dict"].items()})
    if hasattr(gen, "invalidate"):
        gen.invalidate()
    if opt is not None:
        if opt.flat.device != next(gen.parameters()).device:
            raise _lib.ResrError("optimizer and generator live on different devices")
        # load_state_dict copied into the parameter views of the flat master vector: nothing else to move
        if ckpt.get("optimizer"):
            load_optimizer_state_dict(gen, opt, ckpt["optimizer"])
        if ckpt.get("ema_shadow"):
            for name, pos, n, _ in _spans(gen):
                opt.shadow[pos:pos + n].copy_(ckpt["ema_shadow"][name].reshape(-1))
        else:  # a reference file: EMA.register() semantics, the shadow restarts from the loaded weights
            opt.shadow.copy_(opt.flat)
    return ckpt
