"""resr_b200 — sm_100a (B200) implementation of the two hot paths of Lornatang/Real_ESRGAN-PyTorch.

Host-side mirror of the reference's Python API for those paths (same names, arguments and error behaviour):
  model.Generator / ResidualDenseBlock / ResidualResidualDenseBlock      (/root/reference/model.py)
  imgproc.filter2d_torch / USMSharp / DiffJPEG / random_add_*_noise_torch / random_crop (/root/reference/imgproc.py)
  plan.synth_plan / canonical_plan_s0: host decisions + host-drawn tensors of train_realesrnet.py:267-377
  iqa.NIQE                                                               (/root/reference/image_quality_assessment.py:1001-1033)
Everything dispatches through ctypes into the C ABI of lib/libresr.so (include/resr.h). There is no CPU
fallback: importing works anywhere, computing needs the built library and a B200.
"""
from . import _lib  # noqa: F401
from . import autograd  # noqa: F401
from . import checkpoint  # noqa: F401
from . import compat  # noqa: F401
from . import imgproc  # noqa: F401
from . import iqa  # noqa: F401
from . import model  # noqa: F401
from . import optim  # noqa: F401
from . import plan  # noqa: F401

from .compat import patch_reference  # noqa: F401,E402

__all__ = ["_lib", "autograd", "checkpoint", "compat", "imgproc", "iqa", "model", "optim", "plan", "patch_reference"]
