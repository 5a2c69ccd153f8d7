"""Import the unmodified reference from /root/reference (build container only; absent on the GPU box).

Shims, all outside the reference tree:
  1. torchvision.transforms.functional_tensor was removed in torchvision >= 0.17; imgproc.py:27 imports
     rgb_to_grayscale from it.
  2. imgproc.filter2d_torch uses .view on possibly non-contiguous input (imgproc.py:1109,1116) -> wrap with
     .contiguous() (numerically identical).
"""
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model.py"))


def load():
    """Returns (model, imgproc, config) reference modules."""
    if not available():
        raise RuntimeError("/root/reference is not present (it only exists in the build container)")
    import torchvision.transforms._functional_tensor as _ft
    name = "torchvision.transforms.functional_tensor"
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.rgb_to_grayscale = _ft.rgb_to_grayscale
        sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    saved = {k: sys.modules.pop(k) for k in ("model", "imgproc", "config") if k in sys.modules
             and not getattr(sys.modules[k], "__file__", "").startswith(REF)}
    try:
        import config as ref_config
        import imgproc as ref_imgproc
        import model as ref_model
    finally:
        sys.modules.update(saved)
    import torch
    ref_config.device = torch.device("cpu")
    if not getattr(ref_imgproc, "_resr_shimmed", False):
        orig = ref_imgproc.filter2d_torch
        ref_imgproc.filter2d_torch = lambda im, k: orig(im.contiguous(), k)
        ref_imgproc._resr_shimmed = True
    return ref_model, ref_imgproc, ref_config
