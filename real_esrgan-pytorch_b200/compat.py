"""Switching an unmodified reference script to the native hot paths without editing its call sites.

The reference scripts bind three names at import time: `import model`, `import imgproc` and
`from torch.nn import functional as F` (train_realesrnet.py:14-30). `patch_reference(namespace)` rebinds them in a
module / dict namespace to the mirrors of this package, so that `model.Generator(...)`, `imgproc.filter2d_torch(...)`
and the four `F.interpolate(...)` calls of the degradation block (train_realesrnet.py:288, 326, 349, 366) dispatch into
libresr.so with NO edit of the calling lines.
"""
import types

import torch
import torch.nn.functional as _F

from . import imgproc as _imgproc
from . import model as _model

_NATIVE_MODES = ("area", "bilinear", "bicubic")


class _Functional(types.ModuleType):
    """`torch.nn.functional` with `interpolate` routed to the native resize for the calls the degradation block makes
    (CUDA fp32 NCHW image, mode area / bilinear / bicubic, no align_corners / antialias request); every other call and
    every other attribute falls through to torch."""

    def __init__(self):
        super().__init__("resr_b200.functional")

    def __getattr__(self, name):
        return getattr(_F, name)

    @staticmethod
    def interpolate(input, size=None, scale_factor=None, mode="nearest", align_corners=None, recompute_scale_factor=None,
                    antialias=False):
        native = (torch.is_tensor(input) and input.is_cuda and input.dim() == 4 and input.dtype == torch.float32
                  and mode in _NATIVE_MODES and not align_corners and not antialias and not recompute_scale_factor
                  and not (torch.is_grad_enabled() and input.requires_grad))
        if native:
            return _imgproc.interpolate(input, size=size, scale_factor=scale_factor, mode=mode)
        return _F.interpolate(input, size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners,
                              recompute_scale_factor=recompute_scale_factor, antialias=antialias)


functional = _Functional()


def patch_reference(namespace, functional_name: str = "F"):
    """Rebinds `model`, `imgproc` and `F` inside `namespace` (a module object such as the imported
    `train_realesrnet`, or a dict such as `globals()`) to the native mirrors. Returns the names it rebound."""
    get = namespace.get if isinstance(namespace, dict) else (lambda k, d=None: getattr(namespace, k, d))

    def put(k, v):
        if isinstance(namespace, dict):
            namespace[k] = v
        else:
            setattr(namespace, k, v)

    done = []
    for name, repl in (("model", _model), ("imgproc", _imgproc), (functional_name, functional)):
        if get(name) is not None:
            put(name, repl)
            done.append(name)
    return done
