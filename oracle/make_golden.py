"""Generates tests/golden/*.npz from the UNMODIFIED reference in /root/reference (build container only).

    python -m oracle.make_golden            # from the repo root

Versions used for the committed files: torch 2.11.0+cu128 (CPU), torchvision 0.26.0, cv2 4.13.0, scipy 1.18.1,
numpy 2.3.5.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_generator(ref_model):
    out = {}
    for tag, seed, shape in (("a", 0, (1, 3, 32, 32)), ("b", 5, (2, 3, 16, 24))):
        torch.manual_seed(seed)
        g = ref_model.Generator(3, 3, 4).eval()
        torch.manual_seed(1000 + seed)
        x = torch.rand(*shape)
        with torch.no_grad():
            y = g(x)
        out[f"{tag}_seed"] = np.int64(seed)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = y.numpy()
    # cfg1 of BASELINE.json: seed 0 weights, x = torch.rand(1,3,128,128) drawn after manual_seed(0); every 4th pixel
    torch.manual_seed(0)
    g = ref_model.Generator(3, 3, 4).eval()
    torch.manual_seed(0)
    x = torch.rand(1, 3, 128, 128)
    with torch.no_grad():
        y = g(x)
    out["cfg1_y_sub"] = y[:, :, ::4, ::4].contiguous().numpy()
    out["cfg1_y_mean"] = np.float64(y.double().mean().item())
    np.savez_compressed(os.path.join(OUT, "generator.npz"), **out)
    print("generator.npz written")


def make_kernels(ref_imgproc, ref_config):
    """Seeded calls of the reference kernel generators (imgproc.py:492-603); tests replay the same seeds."""
    import math
    import random
    P = ref_config.degradation_model_parameters_dict
    z = {}
    for seed in range(16):
        ks = P["gaussian_kernel_range"][seed % 8]
        random.seed(seed)
        np.random.seed(seed)
        z[f"mixed_{seed}"] = ref_imgproc.random_mixed_kernels(
            P["gaussian_kernel_type"], P["gaussian_kernel_probability1"], ks, P["gaussian_sigma_range1"],
            P["gaussian_sigma_range1"], [-math.pi, math.pi], P["generalized_kernel_beta_range1"],
            P["plateau_kernel_beta_range1"], noise_range=None)
        z[f"mixed_{seed}_ks"] = np.int64(ks)
    for i, (om, ks, pad) in enumerate([(1.2, 7, 0), (2.9, 13, 21), (0.7, 21, 21), (3.1, 9, 21)]):
        z[f"sinc_{i}"] = ref_imgproc.generate_sinc_kernel(om, ks, padding=pad)
        z[f"sinc_{i}_args"] = np.array([om, ks, pad])
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **z)
    print("kernels.npz written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref_model, ref_imgproc, ref_config = refshim.load()
    which = sys.argv[1:] or ["generator", "degrade", "kernels"]
    if "kernels" in which:
        make_kernels(ref_imgproc, ref_config)
    if "generator" in which:
        make_generator(ref_model)
    if "degrade" in which:
        from oracle import make_golden_degrade
        make_golden_degrade.make(ref_imgproc, ref_config, OUT)
