"""Degradation plans (host logic of the second-order degradation path).

A *plan* is every host-side random decision and every host-drawn random tensor of one execution of the reference
degradation block (train_realesrnet.py:267-377), so that `imgproc.degrade_batch` / `imgproc.DegradePipeline` can run the
whole block with no per-sample host synchronisation ("the same blur kernels and noise tensors fed from the host",
BASELINE.json north_star). Layout (dict): blur1, resize1{mode,out_h,out_w,scale}, noise1{type,sigma|scale,gray,
noise_color|samples_color,noise_gray|samples_gray}, jpeg1_quality, blur2, resize2{...}, noise2{...}, final_order
(0: resize->sinc->jpeg, 1: jpeg->resize->sinc), resize3{...}, jpeg2_quality, crop{hr_top,hr_left,image_size,upscale}.

  synth_plan(...)         draws a plan with the reference's distributions (config.py:41-62,
                          train_realesrnet.py:275-371) from a numpy Generator.
  canonical_plan_s0(...)  the canonical plan S0 of SURVEY.md §8d used for roofline accounting.
"""
import numpy as np

AREA, BILINEAR, BICUBIC = 0, 1, 2  # resize mode codes of resr_resize (include/resr.h)

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



def interp_out_size(in_size, scale_factor):
    """F.interpolate(scale_factor=s): floor(float(in) * s) (torch/nn/functional.py)."""
    return int(np.floor(float(in_size) * scale_factor))


def _resize_decision(rng, probs, lo, hi):
    t = rng.choice(3, p=probs)
    if t == 0:
        return float(rng.uniform(1, hi))
    if t == 1:
        return float(rng.uniform(lo, 1))
    return 1.0


def synth_plan(batch, hr_h, hr_w, seed, image_size=256, upscale=4, device_noise=False):
    """Draws a plan with the reference's probabilities/ranges (config.py:41-62, train_realesrnet.py:275-371) from a
    numpy Generator. Gaussian noise fields are drawn here too (numpy) so that a plan is self-contained and replayable by
    the oracle; `device_noise=True` leaves them out and the noise kernel draws them (Philox). Poisson draws depend on the
    image and are always made on the device unless samples are attached to the plan."""
    rng = np.random.default_rng(seed)
    plan = {"blur1": int(rng.uniform() <= 1.0)}
    s = _resize_decision(rng, [0.2, 0.7, 0.1], 0.15, 1.5)
    h1, w1 = interp_out_size(hr_h, s), interp_out_size(hr_w, s)
    plan["resize1"] = {"mode": int(rng.integers(3)), "out_h": h1, "out_w": w1, "scale": s}

    def noise(h, w, sig_rng, sc_rng):
        gray = (rng.uniform(size=batch) < 0.4).astype(np.float32)
        if rng.uniform() < 0.5:
            p = {"type": "gaussian", "sigma": rng.uniform(*sig_rng, size=batch).astype(np.float32), "gray": gray}
            if device_noise:  # the normal deviates are drawn inside the noise kernel (Philox, seed below)
                p["noise_color"], p["seed"] = None, int(rng.integers(1, 2 ** 31))
            else:             # host-drawn fields travel with the plan (replayable by the oracle)
                p["noise_color"] = rng.standard_normal((batch, 3, h, w), dtype=np.float32)
                if gray.sum() > 0:
                    p["noise_gray"] = rng.standard_normal((h, w), dtype=np.float32)
        else:
            # Poisson draws depend on the image: made on the device (sampler "device": inside the fused noise kernel,
            # Philox seed below; "torch": torch.poisson on the rate tensors), or host-fed through samples_color / _gray
            p = {"type": "poisson", "scale": rng.uniform(*sc_rng, size=batch).astype(np.float32), "gray": gray,
                 "samples_color": None, "sampler": "device", "seed": int(rng.integers(1, 2 ** 31))}
        return p

    plan["noise1"] = noise(h1, w1, (1, 30), (0.05, 3))
    plan["jpeg1_quality"] = rng.uniform(30, 95, size=batch).astype(np.float32)
    plan["blur2"] = int(rng.uniform() < 0.8)
    s2 = _resize_decision(rng, [0.3, 0.4, 0.3], 0.3, 1.2)
    h2, w2 = int(hr_h / upscale * s2), int(hr_w / upscale * s2)
    plan["resize2"] = {"mode": int(rng.integers(3)), "out_h": h2, "out_w": w2, "scale": None}
    plan["noise2"] = noise(h2, w2, (1, 25), (0.05, 2.5))
    plan["final_order"] = int(not (rng.uniform() < 0.5))
    plan["resize3"] = {"mode": int(rng.integers(3)), "out_h": hr_h // upscale, "out_w": hr_w // upscale, "scale": None}
    plan["jpeg2_quality"] = rng.uniform(30, 95, size=batch).astype(np.float32)
    plan["crop"] = {"hr_top": int(rng.integers(0, hr_h - image_size + 1)),
                    "hr_left": int(rng.integers(0, hr_w - image_size + 1)), "image_size": image_size, "upscale": upscale}
    return plan


def canonical_plan_s0(batch, hr_h=256, hr_w=256, seed=0):
    """SURVEY.md §8d canonical plan S0: blur1; bicubic x0.5; Gaussian noise with a gray mix; JPEG; blur2; bilinear ->
    H/4; Poisson noise; branch A (area resize (identity), sinc, JPEG); round; crop offset 0."""
    rng = np.random.default_rng(seed)
    h1, w1 = hr_h // 2, hr_w // 2
    h2, w2 = hr_h // 4, hr_w // 4
    gray = np.zeros(batch, np.float32)
    gray[::3] = 1
    return {
        "blur1": 1,
        "resize1": {"mode": BICUBIC, "out_h": h1, "out_w": w1, "scale": 0.5},
        "noise1": {"type": "gaussian", "sigma": rng.uniform(1, 30, size=batch).astype(np.float32), "gray": gray,
                   "noise_color": rng.standard_normal((batch, 3, h1, w1), dtype=np.float32),
                   "noise_gray": rng.standard_normal((h1, w1), dtype=np.float32)},
        "jpeg1_quality": rng.uniform(30, 95, size=batch).astype(np.float32),
        "blur2": 1,
        "resize2": {"mode": BILINEAR, "out_h": h2, "out_w": w2, "scale": None},
        "noise2": {"type": "poisson", "scale": rng.uniform(0.05, 2.5, size=batch).astype(np.float32), "gray": gray,
                   "samples_color": None, "sampler": "device", "seed": int(rng.integers(1, 2 ** 31))},
        "final_order": 0,
        "resize3": {"mode": AREA, "out_h": h2, "out_w": w2, "scale": None},
        "jpeg2_quality": rng.uniform(30, 95, size=batch).astype(np.float32),
        "crop": {"hr_top": 0, "hr_left": 0, "image_size": min(256, hr_h), "upscale": 4},
    }
