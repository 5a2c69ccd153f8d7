"""Oracle for hot path 1: plain torch fp32 restatement of the reference RRDBNet x4 forward. TEST INFRASTRUCTURE ONLY.

Follows /root/reference/model.py:
  ResidualDenseBlock.forward            model.py:87-98
  ResidualResidualDenseBlock.forward    model.py:123-132
  Generator._forward_impl               model.py:255-272
Takes a reference-format state_dict (model.py:223-252 key names). Pinned by tests/golden/generator_*.npz.
"""
import torch
import torch.nn.functional as F


def _conv(x, sd, name):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)


def rdb_forward(x, sd, prefix):
    feats = [x]
    for k in range(1, 5):  # model.py:90-93: dense concat, LeakyReLU(0.2)
        feats.append(F.leaky_relu(_conv(torch.cat(feats, 1), sd, f"{prefix}.conv{k}"), 0.2))
    out5 = _conv(torch.cat(feats, 1), sd, f"{prefix}.conv5")  # model.py:94
    return out5 * 0.2 + x  # model.py:95-96


def rrdb_forward(x, sd, prefix):
    out = x
    for j in (1, 2, 3):  # model.py:126-128
        out = rdb_forward(out, sd, f"{prefix}.rdb{j}")
    return out * 0.2 + x  # model.py:129-130


@torch.no_grad()
def generator_forward(x: torch.Tensor, sd: dict, num_rrdb: int = 23) -> torch.Tensor:
    """x: [N,3,H,W] fp32 in [0,1] -> [N,3,4H,4W] fp32 in [0,1]."""
    sd = {k: v.detach().to(torch.float32) for k, v in sd.items()}
    return _forward(x, sd, num_rrdb)


def l1_loss_and_grads(x: torch.Tensor, hr: torch.Tensor, sd: dict):
    """Reference training-step core under fp32 autograd (train_realesrnet.py:383-388 without AMP): returns
    (loss, {name: grad}) of loss = nn.L1Loss()(G(x), hr)."""
    params = {k: v.detach().clone().to(torch.float32).requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        sr = _forward(x, params, 23)
        loss = torch.nn.functional.l1_loss(sr, hr)
        loss.backward()
    return loss.detach(), {k: p.grad for k, p in params.items()}, sr.detach()


def _forward(x, sd, num_rrdb=23):
    out1 = _conv(x, sd, "conv1")  # model.py:258 (PixelUnshuffle(1) is the identity at x4, model.py:257)
    out = out1
    for i in range(num_rrdb):  # model.py:259
        out = rrdb_forward(out, sd, f"trunk.{i}")
    out = out1 + _conv(out, sd, "conv2")  # model.py:260-262
    out = F.leaky_relu(_conv(F.interpolate(out, scale_factor=2, mode="nearest"), sd, "upsampling1.0"), 0.2)  # :264
    out = F.leaky_relu(_conv(F.interpolate(out, scale_factor=2, mode="nearest"), sd, "upsampling2.0"), 0.2)  # :265
    out = F.leaky_relu(_conv(out, sd, "conv3.0"), 0.2)  # :267
    out = _conv(out, sd, "conv4")  # :268
    return torch.clamp(out, 0.0, 1.0)  # :270 (clamp_ in the reference)


def random_state_dict(seed: int) -> dict:
    """Random-init weights exactly as the reference constructor leaves them (model.py:100-106 + PyTorch defaults),
    drawn in the reference's RNG order, WITHOUT importing the reference: conv default init, then per-RDB
    kaiming_normal_ * 0.1 and zero bias."""
    from torch import nn
    torch.manual_seed(seed)
    sd = {}

    def conv(name, cin, cout):
        m = nn.Conv2d(cin, cout, 3, 1, 1)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.data, m.bias.data
        return m

    conv("conv1", 3, 64)
    for i in range(23):
        for j in (1, 2, 3):
            ms = [conv(f"trunk.{i}.rdb{j}.conv{k + 1}", 64 + 32 * k, 32 if k < 4 else 64) for k in range(5)]
            for m in ms:
                nn.init.kaiming_normal_(m.weight)
                m.weight.data *= 0.1
                nn.init.constant_(m.bias, 0)
    conv("conv2", 64, 64)
    conv("upsampling1.0", 64, 64)
    conv("upsampling2.0", 64, 64)
    conv("conv3.0", 64, 64)
    conv("conv4", 64, 3)
    return sd
