"""Degradation plans. TEST INFRASTRUCTURE ONLY.

A *plan* is every host-side random decision and every random tensor of one execution of the reference degradation
block (train_realesrnet.py:267-377), captured so that the oracle and the CUDA path can replay exactly the same
degradation ("the same blur kernels and noise tensors fed from the host", BASELINE.json north_star).

  record_reference_plan(...)  runs the UNMODIFIED reference block (build container only) and records the plan,
                              every intermediate image and the final (lr, hr).
  synth_plan / canonical_plan_s0 are re-exported from the product's host module (real_esrgan-pytorch_b200/plan.py).
Plan layout (dict): blur1, resize1{mode,out_h,out_w,scale}, noise1{type,sigma|scale,gray,noise_color|samples_color,
noise_gray|samples_gray}, jpeg1_quality, blur2, resize2{...}, noise2{...}, final_order (0: resize->sinc->jpeg,
1: jpeg->resize->sinc), resize3{...}, jpeg2_quality, crop{hr_top,hr_left,image_size,upscale}.
"""
import random
import textwrap

import numpy as np

from . import degrade as od

# config.py:20-39 restated (the reference tree is not available on the GPU box)
DEGRADATION_MODEL_PARAMETERS = {  # config.py:20-39 (restated so the tests do not need the reference tree)
    "sinc_kernel_size": 21, "gaussian_kernel_range": [7, 9, 11, 13, 15, 17, 19, 21],
    "gaussian_kernel_type": ["isotropic", "anisotropic", "generalized_isotropic", "generalized_anisotropic",
                             "plateau_isotropic", "plateau_anisotropic"],
    "gaussian_kernel_probability1": [0.45, 0.25, 0.12, 0.03, 0.12, 0.03], "sinc_kernel_probability1": 0.1,
    "gaussian_sigma_range1": [0.2, 3], "generalized_kernel_beta_range1": [0.5, 4], "plateau_kernel_beta_range1": [1, 2],
    "gaussian_kernel_probability2": [0.45, 0.25, 0.12, 0.03, 0.12, 0.03], "sinc_kernel_probability2": 0.1,
    "gaussian_sigma_range2": [0.2, 1.5], "generalized_kernel_beta_range2": [0.5, 4], "plateau_kernel_beta_range2": [1, 2],
    "sinc_kernel_probability3": 0.8,
}

BLOCK_FIRST_LINE, BLOCK_LAST_LINE = 267, 377  # train_realesrnet.py, 1-based inclusive


def record_reference_plan(hr, kernel1, kernel2, sinc_kernel, seed):
    """hr/kernels: torch CPU tensors. Returns (plan, stages, lr, hr_crop); stages = [(name, input, output)] in order."""
    import torch
    import torch.nn.functional as F

    from . import refshim
    _, imgproc, config = refshim.load()
    src = open(refshim.REF + "/train_realesrnet.py").read().splitlines()[BLOCK_FIRST_LINE - 1:BLOCK_LAST_LINE]
    src = [ln for ln in src if "model.zero_grad" not in ln]
    code = textwrap.dedent("\n".join(src))

    events = []
    draws = []  # torch.rand / randn / poisson outputs inside the current noise call
    state = {"in_usm": False, "in_noise": False, "randint": []}

    usm = imgproc.USMSharp(50, 0)
    jpeger = imgproc.DiffJPEG(False)
    orig = {"filter2d": imgproc.filter2d_torch, "gauss": imgproc.random_add_gaussian_noise_torch,
            "poisson": imgproc.random_add_poisson_noise_torch, "interp": F.interpolate, "rand": torch.rand,
            "randn": torch.randn, "tpoisson": torch.poisson, "randint": random.randint, "crop": imgproc.random_crop}

    def usm_wrap(x, weight, threshold):
        state["in_usm"] = True
        out = usm(x, weight, threshold)
        state["in_usm"] = False
        events.append(("usm", {}, x.clone(), out.clone()))
        return out

    def filter_wrap(image, kernel):
        out = orig["filter2d"](image, kernel)
        if not state["in_usm"]:
            which = "kernel1" if kernel is kernel1 else "kernel2" if kernel is kernel2 else "sinc"
            events.append(("filter2d", {"kernel": which}, image.clone(), out.clone()))
        return out

    def interp_wrap(x, size=None, scale_factor=None, mode="nearest", **kw):
        out = orig["interp"](x, size=size, scale_factor=scale_factor, mode=mode, **kw)
        events.append(("resize", {"mode": od.MODE_ID[mode], "out_h": out.shape[2], "out_w": out.shape[3],
                                  "scale": None if scale_factor is None else float(scale_factor)}, x.clone(), out.clone()))
        return out

    def draw_wrap(name):
        def f(*a, **k):
            t = orig[name](*a, **k)
            if state["in_noise"]:
                draws.append((name, t.clone()))
            return t
        return f

    def noise_wrap(kind):
        def f(image, **kw):
            draws.clear()
            state["in_noise"] = True
            out = orig[kind](image, **kw)
            state["in_noise"] = False
            d = list(draws)
            info = {"type": "gaussian" if kind == "gauss" else "poisson"}
            rng = kw["sigma_range"] if kind == "gauss" else kw["scale_range"]
            first = d[0][1] * (rng[1] - rng[0]) + rng[0]  # imgproc.py:983-985 / 1005-1007
            info["sigma" if kind == "gauss" else "scale"] = first.numpy().astype(np.float32)
            info["gray"] = (d[1][1] < kw["gray_prob"]).float().numpy()
            rest = d[2:]
            if kind == "gauss":
                if len(rest) == 2:
                    info["noise_gray"] = rest[0][1].numpy()
                info["noise_color"] = rest[-1][1].numpy()
            else:
                if len(rest) == 2:
                    info["samples_gray"] = rest[0][1].numpy()
                info["samples_color"] = rest[-1][1].numpy()
            events.append(("noise", info, image.clone(), out.clone()))
            return out
        return f

    def jpeg_wrap(x, quality):
        q = quality.clone()
        out = jpeger(x, quality)
        events.append(("jpeg", {"quality": q.numpy().astype(np.float32)}, x.clone(), out.detach().clone()))
        return out.detach()

    def randint_wrap(a, b):
        v = orig["randint"](a, b)
        state["randint"].append(v)
        return v

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    ns = {"hr": hr, "kernel1": kernel1, "kernel2": kernel2, "sinc_kernel": sinc_kernel, "usm_sharpener": usm_wrap,
          "jpeg_operation": jpeg_wrap, "imgproc": imgproc, "config": config, "np": np, "random": random,
          "torch": torch, "F": F}
    try:
        imgproc.filter2d_torch = filter_wrap
        imgproc.random_add_gaussian_noise_torch = noise_wrap("gauss")
        imgproc.random_add_poisson_noise_torch = noise_wrap("poisson")
        F.interpolate = interp_wrap
        torch.rand, torch.randn, torch.poisson = draw_wrap("rand"), draw_wrap("randn"), draw_wrap("tpoisson")
        random.randint = randint_wrap
        exec(code, ns)
    finally:
        imgproc.filter2d_torch = orig["filter2d"]
        imgproc.random_add_gaussian_noise_torch = orig["gauss"]
        imgproc.random_add_poisson_noise_torch = orig["poisson"]
        F.interpolate = orig["interp"]
        torch.rand, torch.randn, torch.poisson = orig["rand"], orig["randn"], orig["tpoisson"]
        random.randint = orig["randint"]

    # ---- event list -> plan
    names = [e[0] for e in events]
    plan = {}
    it = iter(events)
    ev = next(it)
    assert ev[0] == "usm"
    stages = [("usm", ev[2].numpy(), ev[3].numpy())]
    ev = next(it)
    plan["blur1"] = int(ev[0] == "filter2d")
    if plan["blur1"]:
        stages.append(("blur1", ev[2].numpy(), ev[3].numpy()))
        ev = next(it)
    assert ev[0] == "resize"
    plan["resize1"] = ev[1]
    stages.append(("resize1", ev[2].numpy(), ev[3].numpy()))
    ev = next(it)
    assert ev[0] == "noise"
    plan["noise1"] = ev[1]
    stages.append(("noise1", ev[2].numpy(), ev[3].numpy()))
    ev = next(it)
    assert ev[0] == "jpeg"
    plan["jpeg1_quality"] = ev[1]["quality"]
    stages.append(("jpeg1", ev[2].numpy(), ev[3].numpy()))
    ev = next(it)
    plan["blur2"] = int(ev[0] == "filter2d")
    if plan["blur2"]:
        stages.append(("blur2", ev[2].numpy(), ev[3].numpy()))
        ev = next(it)
    assert ev[0] == "resize"
    plan["resize2"] = ev[1]
    stages.append(("resize2", ev[2].numpy(), ev[3].numpy()))
    ev = next(it)
    assert ev[0] == "noise"
    plan["noise2"] = ev[1]
    stages.append(("noise2", ev[2].numpy(), ev[3].numpy()))
    tail = [next(it), next(it), next(it)]
    tn = [t[0] for t in tail]
    if tn == ["resize", "filter2d", "jpeg"]:
        plan["final_order"] = 0
        r3, sc, jp = tail
    else:
        assert tn == ["jpeg", "resize", "filter2d"], names
        plan["final_order"] = 1
        jp, r3, sc = tail
    plan["resize3"] = r3[1]
    plan["jpeg2_quality"] = jp[1]["quality"]
    for t, nm in sorted(((tail.index(r3), "resize3"), (tail.index(sc), "sinc"), (tail.index(jp), "jpeg2"))):
        stages.append((nm, tail[t][2].numpy(), tail[t][3].numpy()))
    top, left = state["randint"][-2:]
    plan["crop"] = {"hr_top": int(top), "hr_left": int(left), "image_size": int(config.image_size),
                    "upscale": int(config.upscale_factor)}
    return plan, stages, ns["lr"].detach().numpy(), ns["hr"].detach().numpy()


# ------------------------------------------------------------------------------------------------------------------


# Plan synthesis without the reference is host logic of the product (resr_b200.plan); the oracle re-exports it so the
# tests can keep one import.
import importlib.util as _ilu
import os as _os

_spec = _ilu.spec_from_file_location("_resr_plan", _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                                                   "real_esrgan-pytorch_b200", "plan.py"))
_plan = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_plan)
synth_plan = _plan.synth_plan
canonical_plan_s0 = _plan.canonical_plan_s0
