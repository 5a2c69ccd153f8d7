"""Host-side mirror of the reference degradation API (/root/reference/imgproc.py) — same names, arguments and error
behaviour — dispatching into the C ABI (include/resr.h). CUDA fp32 tensors in, new CUDA tensors out; random numbers
come from the global torch / `random` generators in the reference's draw order. No CPU / PyTorch fallback.

  filter2d_torch                     imgproc.py:1089-1121
  USMSharp                           imgproc.py:1514-1537
  DiffJPEG                           imgproc.py:1462-1494
  random_add_gaussian_noise_torch    imgproc.py:1029-1057 (-> :919-940 -> :829-863)
  random_add_poisson_noise_torch     imgproc.py:1060-1086 (-> :943-964 -> :866-916)
  random_crop                        imgproc.py:1894-1934
  interpolate                        torch.nn.functional.interpolate as called by train_realesrnet.py:288,326,349,366
  degrade_batch                      the whole block train_realesrnet.py:267-377 driven by a plan
"""
import ctypes
import math
import random

import torch
from torch import nn

from . import _lib

__all__ = ["image_to_tensor", "tensor_to_image", "filter2d_torch", "USMSharp", "DiffJPEG", "random_add_gaussian_noise_torch",
           "random_add_poisson_noise_torch", "random_crop", "interpolate", "degrade_batch", "degrade_batch_native", "DegradePipeline",
           "plan_to_device"]

_MODES = {"area": 0, "bilinear": 1, "bicubic": 2}


def _on_device(fn):
    """Runs `fn` with the CUDA device of its first tensor argument current, so that the library call, its stream
    (`_lib.stream_ptr()`) and its per-device tables all refer to the device that owns the data."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if torch.is_tensor(a):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def _prep(image: torch.Tensor) -> torch.Tensor:
    if not image.is_cuda:
        raise _lib.ResrError("resr_b200.imgproc runs on CUDA tensors only; there is no CPU path")
    return image.detach().contiguous().float()


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.type, device.index)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def image_to_tensor(image, range_norm: bool, half: bool) -> torch.Tensor:
    """Reference imgproc.py:1540-1567 (torchvision `to_tensor`, restated so that torchvision is not needed): HWC ndarray
    -> CHW tensor; uint8 data is divided by 255, float data keeps its values; optional [-1, 1] scaling and fp16 cast. A
    host-side layout operation exactly as in the reference; the fused device form is `Generator.infer_u8`."""
    import numpy as np
    if not isinstance(image, np.ndarray):
        raise TypeError(f"pic should be ndarray. Got {type(image)}")
    if image.ndim == 2:
        image = image[:, :, None]
    if image.ndim != 3:
        raise ValueError(f"pic should be 2/3 dimensional. Got {image.ndim} dimensions.")
    tensor = torch.from_numpy(np.ascontiguousarray(image.transpose((2, 0, 1))))
    if tensor.dtype == torch.uint8:
        tensor = tensor.to(dtype=torch.get_default_dtype()).div(255)
    if range_norm:
        tensor = tensor.mul(2.0).sub(1.0)
    if half:
        tensor = tensor.half()
    return tensor


def tensor_to_image(tensor: torch.Tensor, range_norm: bool, half: bool):
    """Reference imgproc.py:1570-1596: [1, C, H, W] (or [C, H, W]) tensor in [0, 1] -> HWC uint8 ndarray,
    `mul(255).clamp(0, 255)` then `astype("uint8")` (truncation). CUDA tensors are converted on the device
    (resr_tensor_to_image_u8) so that only H * W * C bytes cross PCIe; host tensors take the reference's own path."""
    if tensor.dim() == 4 and tensor.size(0) != 1:
        raise ValueError("tensor_to_image converts one image: expected [1, C, H, W]")
    if not tensor.is_cuda:
        if range_norm:
            tensor = tensor.add(1.0).div(2.0)
        if half:
            tensor = tensor.half()
        return tensor.squeeze(0).permute(1, 2, 0).mul(255).clamp(0, 255).cpu().numpy().astype("uint8")
    x = tensor.detach().reshape(tensor.shape[-3:]).contiguous().float()
    c, h, w = x.shape
    out = torch.empty((h, w, c), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().resr_tensor_to_image_u8(_lib.ptr(x), _lib.ptr(out), c, h, w, int(bool(range_norm)), int(bool(half)),
                                                      _lib.stream_ptr(x.device)))
    return out.cpu().numpy()


@_on_device
def filter2d_torch(image: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """PyTorch-API twin of cv2.filter2D: reflect pad + cross-correlation (reference imgproc.py:1089-1121)."""
    k = kernel.size(-1)
    b, c, h, w = image.size()
    if k % 2 != 1:
        raise ValueError("Wrong kernel size.")
    x = _prep(image)
    kern = _prep(kernel).view(-1, k, k)
    kb = kern.size(0)
    if kb != 1 and kb != b:
        raise RuntimeError(f"kernel batch {kb} does not match image batch {b}")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().resr_filter2d(_lib.ptr(x), _lib.ptr(kern), _lib.ptr(out), b, c, h, w, k, kb, _lib.stream_ptr()))
    return out


class USMSharp(nn.Module):
    """Unsharp-mask sharpening (reference imgproc.py:1514-1537)."""

    def __init__(self, radius: int, sigma: int) -> None:
        super().__init__()
        if radius % 2 == 0:
            radius += 1
        self.radius = radius
        self.sigma = sigma
        # same buffer as the reference (imgproc.py:1522-1524) so state_dict()/.to() behave alike; the CUDA path
        # rebuilds the separable taps from (radius, sigma)
        s = sigma if sigma > 0 else 0.3 * ((radius - 1) * 0.5 - 1) + 0.8
        xs = torch.arange(radius, dtype=torch.float64) - (radius - 1) * 0.5
        k1 = torch.exp(-(xs * xs) / (2.0 * s * s))
        k1 = k1 / k1.sum()
        self.register_buffer("kernel", torch.outer(k1, k1).float().unsqueeze_(0))

    @_on_device
    def forward(self, x: torch.Tensor, weight: float, threshold: int) -> torch.Tensor:
        if torch.is_grad_enabled() and x.requires_grad:
            # RealESRGAN sharpens the generator output inside its losses (train_realesrgan.py:476-478): differentiable path
            return _USMFn.apply(x, self.radius, int(self.sigma), float(weight), float(threshold))
        return _usm_forward(x, self.radius, int(self.sigma), float(weight), float(threshold))


def _usm_forward(x, radius, sigma, weight, threshold):
    b, c, h, w = x.size()
    xi = _prep(x)
    out = torch.empty_like(xi)
    need = _lib.lib().resr_usm_workspace_bytes(b, c, h, w)
    ws = _workspace(need, xi.device)
    _lib.check(_lib.lib().resr_usm_sharp(_lib.ptr(xi), _lib.ptr(out), b, c, h, w, radius, sigma, weight, threshold, _lib.ptr(ws),
                                         ws.numel(), _lib.stream_ptr()))
    return out


class _USMFn(torch.autograd.Function):
    """USMSharp with its backward (C ABI resr_usm_sharp_backward): autograd of imgproc.py:1526-1535 with the thresholded
    mask treated as the constant it is."""

    @staticmethod
    def forward(ctx, x, radius, sigma, weight, threshold):
        ctx.save_for_backward(x)
        ctx.args = (radius, sigma, weight, threshold)
        with torch.cuda.device(x.device):
            return _usm_forward(x.detach(), radius, sigma, weight, threshold)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        radius, sigma, weight, threshold = ctx.args
        b, c, h, w = x.size()
        xi, gi = _prep(x.detach()), _prep(grad_out)
        out = torch.empty_like(xi)
        with torch.cuda.device(x.device):
            need = _lib.lib().resr_usm_backward_workspace_bytes(b, c, h, w)
            ws = _workspace(need, xi.device)
            _lib.check(_lib.lib().resr_usm_sharp_backward(_lib.ptr(xi), _lib.ptr(gi), _lib.ptr(out), b, c, h, w, radius, sigma, weight,
                                                          threshold, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return out, None, None, None, None


class DiffJPEG(nn.Module):
    """JPEG compress + decompress round trip (reference imgproc.py:1462-1494). differentiable=True (the cubic
    rounding surrogate, imgproc.py:1180-1192) is not on the degradation path and is not implemented."""

    def __init__(self, differentiable: bool) -> None:
        super().__init__()
        if differentiable:
            raise NotImplementedError("resr_b200.DiffJPEG implements differentiable=False (the path the training loops use)")

    @_on_device
    def forward(self, x: torch.Tensor, quality, *, return_coefficients: bool = False, clamp_input: bool = False):
        """clamp_input=True fuses the torch.clamp(out, 0, 1) the training loop applies first (train_realesrnet.py:308)."""
        b, c, h, w = x.size()
        if c != 3:
            raise ValueError("DiffJPEG expects RGB input")
        xi = _prep(x)
        if isinstance(quality, (int, float)):
            q = torch.full((b,), float(quality), dtype=torch.float32, device=xi.device)
            write_back = None
        else:
            write_back = quality
            q = quality.detach().to(device=xi.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(xi)
        factor = torch.empty(b, dtype=torch.float32, device=xi.device)
        qy = qcb = qcr = None
        if return_coefficients:
            hp, wp = (h + 15) // 16 * 16, (w + 15) // 16 * 16
            qy = torch.empty(b, (hp // 8) * (wp // 8), 8, 8, device=xi.device)
            qcb = torch.empty(b, (hp // 16) * (wp // 16), 8, 8, device=xi.device)
            qcr = torch.empty_like(qcb)
        _lib.check(_lib.lib().resr_jpeg(_lib.ptr(xi), _lib.ptr(out), _lib.ptr(q), _lib.ptr(factor), b, h, w, int(clamp_input),
                                        _lib.ptr(qy), _lib.ptr(qcb), _lib.ptr(qcr), _lib.stream_ptr()))
        if write_back is not None:
            # the reference overwrites the caller's quality tensor with the factor (imgproc.py:1474-1479)
            write_back.copy_(factor.to(write_back.device))
        if return_coefficients:
            return out, factor, qy, qcb, qcr
        return out


@_on_device
def gaussian_noise_apply(image, sigma, gray, noise_color, noise_gray, clip=True, rounds=False):
    """Deterministic core of random_add_gaussian_noise_torch: all draws are arguments."""
    b, c, h, w = image.size()
    x = _prep(image)
    out = torch.empty_like(x)
    _lib.check(_lib.lib().resr_gaussian_noise_apply(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(_prep(sigma)), _lib.ptr(None if gray is None else _prep(gray)),
        _lib.ptr(_prep(noise_color)), _lib.ptr(None if noise_gray is None else _prep(noise_gray)), b, c, h, w,
        int(bool(clip)), int(bool(rounds)), _lib.stream_ptr()))
    return out


def random_add_gaussian_noise_torch(image: torch.Tensor, sigma_range: tuple = (0, 1.0), gray_prob: int = 0,
                                    clip: bool = True, rounds: bool = False) -> torch.Tensor:
    """Reference imgproc.py:1029-1057; draws in the reference order: sigma, gray flags, [gray field], colour field."""
    b, _, h, w = image.size()
    kw = dict(dtype=image.dtype, device=image.device)
    sigma = torch.rand(b, **kw) * (sigma_range[1] - sigma_range[0]) + sigma_range[0]
    gray = (torch.rand(b, **kw) < gray_prob).float()
    noise_gray = None
    if torch.sum(gray) > 0:  # the reference branches on this too (imgproc.py:849-855), one host sync
        noise_gray = torch.randn(h, w, **kw)
    noise_color = torch.randn(*image.size(), **kw)
    return gaussian_noise_apply(image, sigma, gray, noise_color, noise_gray, clip, rounds)


_pws_cache = {}


def _poisson_workspace(b: int, device) -> torch.Tensor:
    """Small persistent buffer holding the per-sample level bitmaps / counts / vals between the two Poisson calls."""
    key = (device.type, device.index)
    ws = _pws_cache.get(key)
    need = _lib.lib().resr_poisson_workspace_bytes(b)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 4096), dtype=torch.uint8, device=device)
        _pws_cache[key] = ws
    return ws


@_on_device
def unique_count_u8(image: torch.Tensor, with_gray: bool = True):
    """Per-sample number of distinct u8 levels of the colour image and of its luma (reference imgproc.py:892, 903),
    computed on the device without host syncs. Returns int32 tensors (colour, gray)."""
    b, c, h, w = image.size()
    x = _prep(image)
    cc = torch.empty(b, dtype=torch.int32, device=x.device)
    cg = torch.empty(b, dtype=torch.int32, device=x.device) if with_gray else None
    ws = _poisson_workspace(b, x.device)
    _lib.check(_lib.lib().resr_unique_count_u8(_lib.ptr(x), _lib.ptr(cc), _lib.ptr(cg), b, c, h, w, _lib.ptr(ws),
                                               ws.numel(), _lib.stream_ptr()))
    return cc, cg


@_on_device
def poisson_rates(image: torch.Tensor, with_gray: bool):
    b, c, h, w = image.size()
    x = _prep(image)
    rc = torch.empty_like(x)
    rg = torch.empty(b, 1, h, w, device=x.device) if with_gray else None
    ws = _poisson_workspace(b, x.device)
    _lib.check(_lib.lib().resr_poisson_rates(_lib.ptr(x), _lib.ptr(rc), _lib.ptr(rg), b, c, h, w, _lib.ptr(ws),
                                             ws.numel(), _lib.stream_ptr()))
    return rc, rg


@_on_device
def poisson_noise_apply(image, scale, gray, samples_color, samples_gray, clip=True, rounds=False, reuse_counts=False):
    """Deterministic core of random_add_poisson_noise_torch: all draws are arguments. reuse_counts=True: the unique
    counts of `image` were just computed by poisson_rates(image, ...) (same gray setting) and are reused."""
    b, c, h, w = image.size()
    x = _prep(image)
    out = torch.empty_like(x)
    ws = _poisson_workspace(b, x.device)
    _lib.check(_lib.lib().resr_poisson_noise_apply(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(_prep(scale)), _lib.ptr(None if gray is None else _prep(gray)),
        _lib.ptr(_prep(samples_color)), _lib.ptr(None if samples_gray is None else _prep(samples_gray)), b, c, h, w,
        int(bool(clip)), int(bool(rounds)), _lib.ptr(ws), ws.numel(), int(bool(reuse_counts)), _lib.stream_ptr()))
    return out


_pctr_cache = {}
_gctr_cache = {}


@_on_device
def gaussian_noise_sampled(image, sigma, gray, seed: int, clip=True, rounds=False):
    """Production form of the Gaussian-noise stage for the plan-driven pipeline: the normal deviates are drawn inside the
    kernel (Philox), the gray field is one H x W field shared by the batch (imgproc.py:853-856). `gray=None`: no sample uses
    the gray field. Every call / CUDA-graph replay draws fresh noise for a fixed `seed`."""
    b, c, h, w = image.size()
    x = _prep(image)
    out = torch.empty_like(x)
    key = (x.device.type, x.device.index)
    st = _gctr_cache.get(key)
    if st is None:
        st = torch.zeros(2, dtype=torch.int64, device=x.device)
        _gctr_cache[key] = st
    _lib.check(_lib.lib().resr_gaussian_noise_sampled(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(_prep(sigma)), _lib.ptr(None if gray is None else _prep(gray)), b, c, h, w,
        int(bool(clip)), int(bool(rounds)), int(seed) & (2 ** 64 - 1), _lib.ptr(st), _lib.stream_ptr()))
    return out


@_on_device
def poisson_noise_sampled(image, scale, gray, seed: int, clip=True, rounds=False):
    """Production form of the Poisson stage for the plan-driven pipeline: draws are made inside the kernel (Philox
    counter RNG + exact PTRS / multiplication samplers), so the stage is a memset, the
    presence-bitmap kernel and ONE fused kernel. `gray=None`: no sample takes the luma branch. Every call (and every CUDA
    graph replay) advances a per-device call counter, i.e. draws fresh samples for a fixed `seed`."""
    b, c, h, w = image.size()
    x = _prep(image)
    out = torch.empty_like(x)
    ws = _poisson_workspace(b, x.device)
    key = (x.device.type, x.device.index)
    ctr = _pctr_cache.get(key)
    if ctr is None:
        ctr = torch.zeros(1, dtype=torch.int64, device=x.device)
        _pctr_cache[key] = ctr
    _lib.check(_lib.lib().resr_poisson_noise_sampled(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(_prep(scale)), _lib.ptr(None if gray is None else _prep(gray)), b, c, h, w,
        int(bool(clip)), int(bool(rounds)), int(seed) & (2 ** 64 - 1), _lib.ptr(ctr), _lib.ptr(ws), ws.numel(),
        _lib.stream_ptr()))
    return out


def random_add_poisson_noise_torch(image: torch.Tensor, scale_range: tuple = (0, 1.0), gray_prob: int = 0,
                                   clip: bool = True, rounds: bool = False) -> torch.Tensor:
    """Reference imgproc.py:1060-1086; draws in the reference order: scale, gray flags, [gray Poisson], colour Poisson."""
    b = image.size(0)
    kw = dict(dtype=image.dtype, device=image.device)
    scale = torch.rand(b, **kw) * (scale_range[1] - scale_range[0]) + scale_range[0]
    gray = (torch.rand(b, **kw) < gray_prob).float()
    with_gray = bool(torch.sum(gray) > 0)  # imgproc.py:884-886
    rate_c, rate_g = poisson_rates(image, with_gray)
    samples_gray = torch.poisson(rate_g) if with_gray else None
    samples_color = torch.poisson(rate_c)
    return poisson_noise_apply(image, scale, gray, samples_color, samples_gray, clip, rounds, reuse_counts=True)


@_on_device
def _crop(image, top, left, h_out, w_out, round_to_u8=False):
    b, c, h, w = image.size()
    x = _prep(image)
    out = torch.empty(b, c, h_out, w_out, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().resr_crop(_lib.ptr(x), _lib.ptr(out), b * c, h, w, top, left, h_out, w_out,
                                    int(round_to_u8), _lib.stream_ptr()))
    return out


def random_crop(lr_images: torch.Tensor, hr_images: torch.Tensor, hr_image_size: int, upscale_factor: int):
    """Reference imgproc.py:1894-1934: ONE (top, left) per batch from `random.randint`, LR offset = HR offset // scale."""
    hr_h, hr_w = hr_images[0].size()[1:]
    hr_top = random.randint(0, hr_h - hr_image_size)
    hr_left = random.randint(0, hr_w - hr_image_size)
    lr_size = hr_image_size // upscale_factor
    lr = _crop(lr_images, hr_top // upscale_factor, hr_left // upscale_factor, lr_size, lr_size)
    hr = _crop(hr_images, hr_top, hr_left, hr_image_size, hr_image_size)
    return lr, hr


@_on_device
def interpolate(image: torch.Tensor, size=None, scale_factor=None, mode: str = "bilinear") -> torch.Tensor:
    """torch.nn.functional.interpolate for modes area / bilinear / bicubic (align_corners=False, no antialias) with
    ATen's index rules; scale_factor= uses 1/scale_factor as the coordinate scale, size= uses in/out."""
    if mode not in _MODES:
        raise NotImplementedError(f"mode {mode!r}: the degradation path uses area / bilinear / bicubic")
    b, c, h, w = image.size()
    if (size is None) == (scale_factor is None):
        raise ValueError("only one of size or scale_factor should be defined")
    if size is not None:
        oh, ow = (size, size) if isinstance(size, int) else size
        sh = sw = 0.0
    else:
        sh, sw = (scale_factor, scale_factor) if not isinstance(scale_factor, (tuple, list)) else scale_factor
        oh, ow = int(math.floor(float(h) * sh)), int(math.floor(float(w) * sw))
    x = _prep(image)
    if (oh, ow) == (h, w) and (size is not None or (float(sh) == 1.0 and float(sw) == 1.0)):
        # same-size resize: all three modes return the input bit for bit (`scale_factor=1` of the first
        # resize, the third resize after a "keep" second one) -- no launch
        return x.clone() if x.data_ptr() == image.data_ptr() else x
    out = torch.empty(b, c, oh, ow, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().resr_resize(_lib.ptr(x), _lib.ptr(out), b * c, h, w, oh, ow, _MODES[mode], float(sh), float(sw),
                                      _lib.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------------------------


def _dev(t, device):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


def _noise(x, p):
    dev = x.device
    if p["type"] == "gaussian":
        if p.get("noise_color") is None:  # plan without host-drawn fields: draw inside the kernel
            g_any = p.get("gray_any")
            gray = _dev(p["gray"], dev)
            with_gray = bool(gray.sum() > 0) if g_any is None else bool(g_any)
            return gaussian_noise_sampled(x, _dev(p["sigma"], dev), gray if with_gray else None, int(p.get("seed", 0)))
        return gaussian_noise_apply(x, _dev(p["sigma"], dev), _dev(p["gray"], dev), _dev(p["noise_color"], dev),
                                    _dev(p.get("noise_gray"), dev))
    sc, sg = p.get("samples_color"), p.get("samples_gray")
    gray = _dev(p["gray"], dev)
    if sc is None:  # plan without recorded draws
        g_any = p.get("gray_any")  # host-side decision recorded by plan_to_device (no device sync)
        with_gray = bool(gray.sum() > 0) if g_any is None else bool(g_any)
        if p.get("sampler", "device") == "device":  # draws inside the fused kernel
            return poisson_noise_sampled(x, _dev(p["scale"], dev), gray if with_gray else None, int(p.get("seed", 0)))
        # "torch": sample from the rates with the torch generator (same sampler, three more launches)
        rc, rg = poisson_rates(x, with_gray)
        sg = torch.poisson(rg) if with_gray else None
        sc = torch.poisson(rc)
        return poisson_noise_apply(x, _dev(p["scale"], dev), gray, sc, sg, reuse_counts=True)
    return poisson_noise_apply(x, _dev(p["scale"], dev), gray, _dev(sc, dev), _dev(sg, dev))


def plan_to_device(plan: dict, device) -> dict:
    """Copy of `plan` with every array moved to `device` once (so that replaying it launches no H2D copies) and the
    host-side `gray_any` decisions recorded (so that no step has to synchronise on a device flag)."""
    import numpy as np
    out = {}
    for k, v in plan.items():
        if isinstance(v, dict):
            d = {}
            for kk, vv in v.items():
                if isinstance(vv, np.ndarray):
                    d[kk] = torch.from_numpy(np.ascontiguousarray(vv)).to(device=device, dtype=torch.float32)
                else:
                    d[kk] = vv
            if "gray" in v:
                g = v["gray"]
                d["gray_any"] = bool(g.sum() > 0) if not torch.is_tensor(g) else bool(g.sum().item() > 0)
            out[k] = d
        elif isinstance(v, np.ndarray):
            out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=torch.float32)
        else:
            out[k] = v
    return out


class DegradePipeline:
    """The degradation block for a fixed plan signature (branches, resize modes and sizes) with its whole launch
    sequence captured in ONE CUDA graph: replaying a batch costs a single graph launch instead of ~25 kernel launches.
    Inputs live in static device buffers (`hr`, `kernel1`, `kernel2`, `sinc_kernel` and the tensors inside `plan`);
    update them in place (``pipe.hr.copy_(new_hr)``) and call the pipeline again."""

    def __init__(self, hr, kernel1, kernel2, sinc_kernel, plan, use_graph: bool = True, u8_images: bool = False):
        dev = hr.device
        self.hr, self.kernel1, self.kernel2, self.sinc_kernel = (t.detach().clone().float().contiguous()
                                                                 for t in (hr, kernel1, kernel2, sinc_kernel))
        self.plan = plan_to_device(plan, dev)
        self.graph = None
        # u8_images: the batch arrives as decoded u8 HWC BGR images (`pipe.images_u8`, with the per-sample augmentation ops in
        # `pipe.augment_ops`); the augmentation gather that turns them into `hr` (dataset.py:67-79) is the graph's first node
        self.images_u8 = self.augment_ops = None
        if u8_images:
            b, _, h, w = self.hr.shape
            self.images_u8 = torch.zeros(b, h, w, 3, dtype=torch.uint8, device=dev)
            self.augment_ops = torch.zeros(b, dtype=torch.int32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside the capture: constant-table uploads, function attributes, workspaces
                self.lr, self.hr_crop = self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.lr, self.hr_crop = self._run()
            self.graph = g
        self.launches = None

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.lr, self.hr_crop = self._run()
        return self.lr, self.hr_crop

    def _run(self):
        if self.images_u8 is not None:
            augment_batch(self.images_u8, self.augment_ops, out=self.hr)
        return degrade_batch(self.hr, self.kernel1, self.kernel2, self.sinc_kernel, self.plan)


_block_mods = None


def _block_modules():
    """USMSharp(50, 0) / DiffJPEG(False) of the degradation block (train_realesrnet.py:231-235), built once: the 51 x 51
    buffer of USMSharp is a host-side outer product that has no business in a per-batch call."""
    global _block_mods
    if _block_mods is None:
        _block_mods = (USMSharp(50, 0), DiffJPEG(False))
    return _block_mods


@_on_device
def degrade_batch(hr: torch.Tensor, kernel1: torch.Tensor, kernel2: torch.Tensor, sinc_kernel: torch.Tensor, plan: dict,
                  stages: list = None):
    """The reference's second-order degradation block (train_realesrnet.py:267-377) with every host decision and
    random tensor taken from `plan` (layout: oracle/plan.py). Returns (lr, hr_crop); lr is detached and on the u8 grid."""
    usm, jpeger = _block_modules()
    dev = hr.device

    def rec(name, t):
        if stages is not None:
            stages.append((name, t))
        return t

    def jpeg(x, q):
        return jpeger(x, _dev(q, dev).clone(), clamp_input=True)

    def resize(x, r):
        if r.get("scale") is not None:
            return interpolate(x, scale_factor=r["scale"], mode=("area", "bilinear", "bicubic")[r["mode"]])
        return interpolate(x, size=(r["out_h"], r["out_w"]), mode=("area", "bilinear", "bicubic")[r["mode"]])

    out = rec("usm", usm(hr, 0.5, 10))
    if plan["blur1"]:
        out = rec("blur1", filter2d_torch(out, kernel1))
    out = rec("resize1", resize(out, plan["resize1"]))
    out = rec("noise1", _noise(out, plan["noise1"]))
    out = rec("jpeg1", jpeg(out, plan["jpeg1_quality"]))
    if plan["blur2"]:
        out = rec("blur2", filter2d_torch(out, kernel2))
    out = rec("resize2", resize(out, plan["resize2"]))
    out = rec("noise2", _noise(out, plan["noise2"]))
    if plan["final_order"] == 0:
        out = rec("resize3", resize(out, plan["resize3"]))
        out = rec("sinc", filter2d_torch(out, sinc_kernel))
        out = rec("jpeg2", jpeg(out, plan["jpeg2_quality"]))
    else:
        out = rec("jpeg2", jpeg(out, plan["jpeg2_quality"]))
        out = rec("resize3", resize(out, plan["resize3"]))
        out = rec("sinc", filter2d_torch(out, sinc_kernel))
    c = plan["crop"]
    ls = c["image_size"] // c["upscale"]
    lr = _crop(out, c["hr_top"] // c["upscale"], c["hr_left"] // c["upscale"], ls, ls, round_to_u8=True)
    hr_c = _crop(hr, c["hr_top"], c["hr_left"], c["image_size"], c["image_size"])
    return lr, hr_c


_rng_state_cache = {}


@_on_device
def degrade_batch_native(hr: torch.Tensor, kernel1: torch.Tensor, kernel2: torch.Tensor, sinc_kernel: torch.Tensor, plan: dict):
    """Same block as `degrade_batch`, sequenced inside the library by ONE C-ABI call (`resr_degrade_batch`): the plan dict
    is flattened into the POD `resr_degrade_plan` (host decisions + device pointers). Returns (lr, hr_crop)."""
    dev = hr.device
    b, c, h, w = hr.size()
    keep = []  # device tensors referenced by raw pointers must outlive the call

    def d(t):
        t = _dev(t, dev)
        keep.append(t)
        return _lib.ptr(t)

    P = _lib.DegradePlan()
    P.batch, P.hr_h, P.hr_w = b, h, w
    P.usm_radius, P.usm_sigma, P.usm_weight, P.usm_threshold = 50, 0, 0.5, 10.0
    P.blur1, P.blur2, P.final_order = int(plan["blur1"]), int(plan["blur2"]), int(plan["final_order"])
    P.kernel_size, P.sinc_batched = int(kernel1.size(-1)), int(sinc_kernel.size(0) != 1)
    for name in ("resize1", "resize2", "resize3"):
        r, spec = plan[name], getattr(P, name)
        spec.mode, spec.out_h, spec.out_w = int(r["mode"]), int(r["out_h"]), int(r["out_w"])
        spec.scale = float(r["scale"]) if r.get("scale") is not None else 0.0
    need_rng = False
    for name in ("noise1", "noise2"):
        n, spec = plan[name], getattr(P, name)
        gray = _dev(n["gray"], dev)
        g_any = n.get("gray_any")
        spec.gray_any = int(bool(gray.sum() > 0) if g_any is None else bool(g_any))
        keep.append(gray)
        spec.gray = _lib.ptr(gray)
        spec.seed = int(n.get("seed", 0))
        if n["type"] == "gaussian":
            spec.type, spec.param = 0, d(n["sigma"])
            if n.get("noise_color") is not None:
                spec.draws_color = d(n["noise_color"])
                spec.draws_gray = d(n["noise_gray"]) if n.get("noise_gray") is not None else None
            else:
                need_rng = True
        else:
            spec.type, spec.param = 1, d(n["scale"])
            if n.get("samples_color") is not None:
                spec.draws_color = d(n["samples_color"])
                spec.draws_gray = d(n["samples_gray"]) if n.get("samples_gray") is not None else None
            else:
                need_rng = True
    P.jpeg1_quality, P.jpeg2_quality = d(plan["jpeg1_quality"]), d(plan["jpeg2_quality"])
    cr = plan["crop"]
    P.crop_top, P.crop_left, P.image_size, P.upscale = int(cr["hr_top"]), int(cr["hr_left"]), int(cr["image_size"]), int(cr["upscale"])
    if need_rng:
        key = (dev.type, dev.index)
        st = _rng_state_cache.get(key)
        if st is None:
            st = torch.zeros(8, dtype=torch.int64, device=dev)
            _rng_state_cache[key] = st
        P.rng_state = _lib.ptr(st)
    x, k1, k2, sk = _prep(hr), _prep(kernel1), _prep(kernel2), _prep(sinc_kernel)
    ls = P.image_size // P.upscale
    lr = torch.empty(b, 3, ls, ls, dtype=torch.float32, device=dev)
    hr_c = torch.empty(b, 3, P.image_size, P.image_size, dtype=torch.float32, device=dev)
    need = _lib.lib().resr_degrade_workspace_bytes(ctypes.byref(P))
    ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 256
    _lib.check(_lib.lib().resr_degrade_batch(ctypes.byref(P), _lib.ptr(x), _lib.ptr(k1), _lib.ptr(k2), _lib.ptr(sk), _lib.ptr(lr),
                                             _lib.ptr(hr_c), wp, need, _lib.stream_ptr()))
    return lr, hr_c


# ------------------------------------------------------------------------------------------------------------------
# Training-image augmentation of the dataset (dataset.py:66-79) for a whole batch of decoded images.

def draw_augment_ops(batch: int, angles=(0, 90, 180, 270), p_hflip: float = 0.5, p_vflip: float = 0.5):
    """The random decisions of dataset.py:71-73 for `batch` images, drawn from Python's `random` in the reference's order
    (per image: random.choice(angles) in random_rotate, random.random() < p in the horizontal, then the vertical flip).
    Returns an int32 CPU tensor of packed ops (bits 0-1 angle index, bit 2 horizontal flip, bit 3 vertical flip)."""
    import random
    if tuple(angles) != (0, 90, 180, 270):
        raise ValueError("the device augmentation implements the reference's angles [0, 90, 180, 270]")
    ops = []
    for _ in range(batch):
        ai = angles.index(random.choice(list(angles)))
        hf = random.random() < p_hflip
        vf = random.random() < p_vflip
        ops.append(ai | (int(hf) << 2) | (int(vf) << 3))
    return torch.tensor(ops, dtype=torch.int32)


@_on_device
def augment_batch(images_u8_bgr: torch.Tensor, ops: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """dataset.py:67-79 on the device: decoded u8 HWC BGR images [b, h, w, 3] -> fp32 RGB tensors [b, 3, h, w] in [0, 1],
    rotated / flipped per sample as `ops` says (see draw_augment_ops). One gather kernel, bit-exact against
    random_rotate + random_*_flip + cvtColor + image_to_tensor of the reference (C ABI resr_augment_batch_u8)."""
    if not images_u8_bgr.is_cuda:
        raise _lib.ResrError("resr_b200.imgproc runs on CUDA tensors only; there is no CPU path")
    if images_u8_bgr.dtype != torch.uint8 or images_u8_bgr.dim() != 4 or images_u8_bgr.size(-1) != 3:
        raise ValueError("expected a uint8 tensor [b, h, w, 3]")
    x = images_u8_bgr.contiguous()
    b, h, w, _ = x.shape
    o = ops.to(device=x.device, dtype=torch.int32).contiguous()
    if o.numel() != b:
        raise ValueError("one op per image")
    if out is None:
        out = torch.empty(b, 3, h, w, dtype=torch.float32, device=x.device)
    elif out.shape != (b, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous fp32 tensor [b, 3, h, w] on the images' device")
    _lib.check(_lib.lib().resr_augment_batch_u8(_lib.ptr(x), _lib.ptr(out), _lib.ptr(o), b, h, w, _lib.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------------------------
# Blur-kernel synthesis (reference imgproc.py:225-603, dataset.py:81-141): random draws on the host in the reference's
# RNG order, arithmetic in float64 on the device (C ABI resr_synthesize_kernels).

_KTYPES = {"gaussian": 0, "generalized": 1, "plateau": 2, "sinc": 3, "delta": 4}


def _synth(params, pad=0, device=None, dtype=torch.float64):
    """params: list of dicts (type, kernel_size, isotropic, sigma_x, sigma_y, theta, beta, cutoff) -> [n, P, P] tensor."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = len(params)
    arr = (_lib.KernelParams * n)()
    for i, p in enumerate(params):
        arr[i].type = _KTYPES[p["type"]]
        arr[i].kernel_size = int(p["kernel_size"])
        arr[i].isotropic = int(bool(p.get("isotropic", True)))
        arr[i].sigma_x = float(p.get("sigma_x", 1.0))
        arr[i].sigma_y = float(p.get("sigma_y", 1.0))
        arr[i].theta = float(p.get("theta", 0.0))
        arr[i].beta = float(p.get("beta", 1.0))
        arr[i].cutoff = float(p.get("cutoff", 1.0))
    P = pad if pad else int(params[0]["kernel_size"])
    out = torch.empty(n, P, P, dtype=dtype, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().resr_synthesize_kernels(arr, n, int(pad), _lib.ptr(out) if dtype == torch.float64 else None,
                                                      _lib.ptr(out) if dtype == torch.float32 else None, _lib.stream_ptr()))
    return out


def generate_sinc_kernel(cutoff: float, kernel_size: int, padding: int = 0):
    """Reference imgproc.py:576-603. Returns a float64 ndarray like the reference."""
    assert kernel_size % 2 == 1, "Kernel size must be an odd number."
    pad = padding if padding > kernel_size else 0
    return _synth([{"type": "sinc", "kernel_size": kernel_size, "cutoff": cutoff}], pad)[0].cpu().numpy()


def _draw_gaussian_family(kind, kernel_size, sigma_x_range, sigma_y_range, rotation_range, beta_range, isotropic):
    """RNG draw order of imgproc.py:330-489 (_random_bivariate_*_kernel)."""
    import numpy as np
    assert kernel_size % 2 == 1, "Kernel size must be an odd number."
    assert sigma_x_range[0] < sigma_x_range[1], "Wrong sigma_x_range."
    sigma_x = np.random.uniform(sigma_x_range[0], sigma_x_range[1])
    if isotropic is False:
        assert sigma_y_range[0] < sigma_y_range[1], "Wrong sigma_y_range."
        assert rotation_range[0] < rotation_range[1], "Wrong rotation_range."
        sigma_y = np.random.uniform(sigma_y_range[0], sigma_y_range[1])
        rotation = np.random.uniform(rotation_range[0], rotation_range[1])
    else:
        sigma_y, rotation = sigma_x, 0
    beta = 1.0
    if kind != "gaussian":
        if np.random.uniform() < 0.5:
            beta = np.random.uniform(beta_range[0], 1)
        else:
            beta = np.random.uniform(1, beta_range[1])
    return {"type": kind, "kernel_size": kernel_size, "isotropic": isotropic, "sigma_x": sigma_x, "sigma_y": sigma_y,
            "theta": rotation, "beta": beta}


def draw_mixed_kernel_params(kernel_type, kernel_prob, kernel_size, sigma_x_range, sigma_y_range, rotation_range,
                             generalized_kernel_beta_range, plateau_kernel_beta_range):
    """The host-side random decisions of random_mixed_kernels (imgproc.py:492-573), without the arithmetic."""
    kt = random.choices(kernel_type, kernel_prob)[0]
    iso = not kt.endswith("anisotropic")
    if kt.startswith("generalized"):
        return _draw_gaussian_family("generalized", kernel_size, sigma_x_range, sigma_y_range, rotation_range,
                                     generalized_kernel_beta_range, iso), True
    if kt.startswith("plateau"):
        return _draw_gaussian_family("plateau", kernel_size, sigma_x_range, sigma_y_range, rotation_range,
                                     plateau_kernel_beta_range, iso), False
    if kt not in ("isotropic", "anisotropic"):
        iso = True  # imgproc.py:566-572: unknown names fall back to the isotropic Gaussian
    return _draw_gaussian_family("gaussian", kernel_size, sigma_x_range, sigma_y_range, rotation_range, None, iso), True


def random_mixed_kernels(kernel_type, kernel_prob, kernel_size, sigma_x_range, sigma_y_range, rotation_range,
                         generalized_kernel_beta_range, plateau_kernel_beta_range, noise_range=None):
    """Reference imgproc.py:492-573. Returns a float64 ndarray like the reference."""
    import numpy as np
    p, noisy = draw_mixed_kernel_params(kernel_type, kernel_prob, kernel_size, sigma_x_range, sigma_y_range, rotation_range,
                                        generalized_kernel_beta_range, plateau_kernel_beta_range)
    k = _synth([p])[0].cpu().numpy()
    if noise_range is not None and noisy:  # multiplicative kernel noise (imgproc.py:367-370); plateau kernels get None
        assert noise_range[0] < noise_range[1], "Wrong noise range."
        k = k * np.random.uniform(noise_range[0], noise_range[1], size=k.shape)
        k = k / np.sum(k)
    return k


_draw_state = {}


def draw_config(parameters: dict) -> "_lib.KernelDrawConfig":
    """config.degradation_model_parameters_dict (config.py:20-39) -> the POD struct of resr_draw_degradation_kernel_params."""
    import numpy as np
    P = parameters
    c = _lib.KernelDrawConfig()
    sizes = list(P["gaussian_kernel_range"])
    c.n_sizes = len(sizes)
    for i, v in enumerate(sizes):
        c.sizes[i] = int(v)
    c.sinc_size_split = int(np.median(sizes))
    c.final_size = int(P["sinc_kernel_size"])
    c.sinc_prob1, c.sinc_prob2, c.sinc_prob3 = (float(P[f"sinc_kernel_probability{i}"]) for i in (1, 2, 3))
    for i in range(6):
        c.prob1[i] = float(P["gaussian_kernel_probability1"][i])
        c.prob2[i] = float(P["gaussian_kernel_probability2"][i])
    for i in range(2):
        c.sigma_range1[i], c.sigma_range2[i] = float(P["gaussian_sigma_range1"][i]), float(P["gaussian_sigma_range2"][i])
        c.gen_beta_range1[i], c.gen_beta_range2[i] = float(P["generalized_kernel_beta_range1"][i]), float(P["generalized_kernel_beta_range2"][i])
        c.plateau_beta_range1[i], c.plateau_beta_range2[i] = float(P["plateau_kernel_beta_range1"][i]), float(P["plateau_kernel_beta_range2"][i])
    return c


def draw_degradation_kernels_device(batch: int, parameters: dict, device=None, seed: int = 0, return_params: bool = False):
    """kernel1, kernel2, sinc_kernel for `batch` samples with NO host loop: the per-sample random
    decisions of dataset.py:81-141 are drawn by a device kernel (Philox, the reference's distributions) and evaluated in
    float64 by the synthesis kernel — two launches per batch whatever its size. Every call draws fresh parameters for a
    fixed `seed`. Returns three [batch, P, P] fp32 CUDA tensors (P = largest kernel size)."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    cfg = draw_config(parameters)
    pad = int(parameters["gaussian_kernel_range"][-1])
    kmax = max(max(parameters["gaussian_kernel_range"]), int(parameters["sinc_kernel_size"]))
    pad = max(pad, kmax)
    key = (device.type, device.index)
    st = _draw_state.get(key)
    if st is None:
        st = torch.zeros(1, dtype=torch.int64, device=device)
        _draw_state[key] = st
    params = torch.empty(batch * 3 * ctypes.sizeof(_lib.KernelParams), dtype=torch.uint8, device=device)
    out = torch.empty(batch * 3, pad, pad, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().resr_draw_degradation_kernel_params(ctypes.byref(cfg), batch, int(seed) & (2 ** 64 - 1), _lib.ptr(st),
                                                                  _lib.ptr(params), _lib.stream_ptr(device)))
        _lib.check(_lib.lib().resr_synthesize_kernels_device(_lib.ptr(params), batch * 3, kmax, pad, None, _lib.ptr(out),
                                                             _lib.stream_ptr(device)))
    out = out.view(batch, 3, pad, pad)
    ks = (out[:, 0].contiguous(), out[:, 1].contiguous(), out[:, 2].contiguous())
    return ks + (params,) if return_params else ks


def synthesize_degradation_kernels(batch: int, parameters: dict, device=None):
    """kernel1, kernel2, sinc_kernel for `batch` samples, sequenced per sample exactly as dataset.py:81-141 does
    (same `random` / `np.random` draw order), synthesised in ONE device launch. Returns three [batch, 21, 21] fp32
    CUDA tensors. (`draw_degradation_kernels_device` also makes the random decisions on the device.)"""
    import numpy as np
    P = parameters
    ps = []
    for _ in range(batch):
        for which in (1, 2):
            ks = random.choice(P["gaussian_kernel_range"])
            if np.random.uniform() < P[f"sinc_kernel_probability{which}"]:
                if ks < int(np.median(P["gaussian_kernel_range"])):
                    om = np.random.uniform(np.pi / 3, np.pi)
                else:
                    om = np.random.uniform(np.pi / 5, np.pi)
                ps.append({"type": "sinc", "kernel_size": ks, "cutoff": om})
            else:
                p, _ = draw_mixed_kernel_params(P["gaussian_kernel_type"], P[f"gaussian_kernel_probability{which}"], ks,
                                                P[f"gaussian_sigma_range{which}"], P[f"gaussian_sigma_range{which}"],
                                                [-math.pi, math.pi], P[f"generalized_kernel_beta_range{which}"],
                                                P[f"plateau_kernel_beta_range{which}"])
                ps.append(p)
        if np.random.uniform() < P["sinc_kernel_probability3"]:
            ks = random.choice(P["gaussian_kernel_range"])
            om = np.random.uniform(np.pi / 3, np.pi)
            ps.append({"type": "sinc", "kernel_size": ks, "cutoff": om})
        else:
            ps.append({"type": "delta", "kernel_size": P["sinc_kernel_size"]})
    pad = P["gaussian_kernel_range"][-1]
    out = _synth(ps, pad, device, torch.float32).view(batch, 3, pad, pad)
    return out[:, 0].contiguous(), out[:, 1].contiguous(), out[:, 2].contiguous()
