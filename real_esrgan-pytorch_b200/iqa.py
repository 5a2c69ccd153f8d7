"""NIQE on the device: mirror of the reference's `NIQE` module (image_quality_assessment.py:1001-1033 -> `_niqe_torch`
:886-998). The per-block feature extraction (Y channel, MSCN coefficients, AGGD fits at two scales, MATLAB-style
half-size resize) runs in libresr.so (`resr_niqe_features`, csrc/niqe.cu, float64 like the reference); the 36-dimensional
Gaussian fit against the pristine statistics is a handful of torch.linalg calls on the same device.

The pristine statistics (`niqe_model.mat`, config.py:72) are a download: pass the path of the .mat file exactly as the
reference does, or the arrays themselves (`mu_prisparam` [36], `cov_prisparam` [36, 36])."""
import torch
from torch import nn

from . import _lib


def _load_pristine(model):
    if isinstance(model, (tuple, list)):
        mu, cov = model
    else:
        import numpy as np
        import scipy.io                                           # image_quality_assessment.py:976-979
        m = scipy.io.loadmat(model)
        mu, cov = np.ravel(m["mu_prisparam"]), m["cov_prisparam"]
    mu = torch.as_tensor(mu, dtype=torch.float64).reshape(-1)
    cov = torch.as_tensor(cov, dtype=torch.float64)
    if mu.numel() != 36 or tuple(cov.shape) != (36, 36):
        raise ValueError("pristine statistics must be a 36-vector and a 36 x 36 matrix")
    return mu, cov


def niqe_features(raw_tensor: torch.Tensor, crop_border: int, block_size: int = 96) -> torch.Tensor:
    """[b, 3, h, w] fp32 CUDA tensor in [0, 1] -> float64 [b, blocks, 36] (C ABI resr_niqe_features)."""
    if not raw_tensor.is_cuda:
        raise _lib.ResrError("resr_b200.iqa runs on CUDA tensors only; there is no CPU path")
    if raw_tensor.dim() != 4 or raw_tensor.size(1) != 3:
        raise ValueError("expected an RGB tensor [b, 3, h, w]")
    x = raw_tensor.detach().contiguous().float()
    b, _, h, w = x.shape
    lib = _lib.lib()
    nb = lib.resr_niqe_num_blocks(h, w, int(crop_border), int(block_size))
    if nb <= 0:
        raise ValueError(f"image {h}x{w} holds no {block_size}x{block_size} block after cropping {crop_border} pixels")
    with torch.cuda.device(x.device):
        need = lib.resr_niqe_workspace_bytes(b, h, w, int(crop_border), int(block_size))
        ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        feat = torch.empty(b, nb, 36, dtype=torch.float64, device=x.device)
        _lib.check(lib.resr_niqe_features(_lib.ptr(x), _lib.ptr(feat), b, h, w, int(crop_border), int(block_size), _lib.ptr(ws), need,
                                          _lib.stream_ptr(x.device)))
    return feat


def niqe_from_features(feat: torch.Tensor, mu_pris: torch.Tensor, cov_pris: torch.Tensor) -> torch.Tensor:
    """The multivariate-Gaussian distance of image_quality_assessment.py:879-884 (nanmean / nancov / pinv) per image."""
    mu_pris, cov_pris = mu_pris.to(feat), cov_pris.to(feat)
    out = []
    for f in feat:                                                 # rows with a NaN are dropped from the covariance (:631-644)
        nan_rows = torch.isnan(f).any(dim=1)
        nan = torch.isnan(f)
        mu = torch.where(nan, torch.zeros_like(f), f).sum(0) / (~nan).to(f).sum(0)
        g = f[~nan_rows]
        g = g - g.mean(0, keepdim=True)
        cov = g.t() @ g / (g.shape[0] - 1)
        inv = torch.linalg.pinv((cov_pris + cov) / 2)
        d = (mu_pris - mu).unsqueeze(0)
        out.append(torch.sqrt(d @ inv @ d.t()).reshape(()))
    return torch.stack(out)


class NIQE(nn.Module):
    """Same constructor and call as the reference (image_quality_assessment.py:1001-1033); block height and width must be
    equal and even here."""

    def __init__(self, crop_border: int, niqe_model_path, block_size_height: int = 96, block_size_width: int = 96) -> None:
        super().__init__()
        if block_size_height != block_size_width:
            raise ValueError("resr_b200.iqa.NIQE implements square blocks (the reference's default 96 x 96)")
        self.crop_border = crop_border
        self.niqe_model_path = niqe_model_path
        self.block_size = block_size_height
        self._pristine = None

    def forward(self, raw_tensor: torch.Tensor) -> torch.Tensor:
        if self._pristine is None:
            self._pristine = _load_pristine(self.niqe_model_path)
        feat = niqe_features(raw_tensor, self.crop_border, self.block_size)
        return niqe_from_features(feat, *self._pristine).squeeze()
