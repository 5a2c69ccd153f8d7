"""Writes tests/golden/niqe.npz from the UNMODIFIED reference (build container only):
    python -m oracle.make_golden_niqe
The reference's pretrained NIQE statistics (niqe_model.mat, config.py:72) are a download that is not in the reference tree,
so the fixture uses SYNTHETIC pristine statistics (a random mean vector and a random symmetric positive-definite
covariance, written to a temporary .mat in the layout image_quality_assessment.py:976-979 reads). Recorded per case: the
RGB input (stored as its u8 levels: the tensor is levels / 255 in fp32), crop_border, the statistics, the [blocks, 36] feature matrix the reference fits its Gaussian to (captured
at the entry of _nanmean_torch) and the NIQE score of image_quality_assessment._niqe_torch."""
import os
import sys
import tempfile

import numpy as np
import torch

from . import refshim


def main():
    import scipy.io
    refshim.load()
    sys.path.insert(0, refshim.REF)
    import image_quality_assessment as iqa
    rng = np.random.default_rng(11)
    mu = rng.normal(0.0, 1.0, 36)
    mu[0::18] += 2.5           # shape parameters of the MSCN fit live around 2-3
    a = rng.normal(0.0, 0.3, (36, 36))
    cov = a @ a.T + 0.05 * np.eye(36)
    out = {"mu_prisparam": mu, "cov_prisparam": cov}
    captured = {}
    orig = iqa._nanmean_torch

    def spy(v, *args, **kwargs):
        captured["distparam"] = v.detach().clone()
        return orig(v, *args, **kwargs)

    iqa._nanmean_torch = spy
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "niqe_model.mat")
        scipy.io.savemat(path, {"mu_prisparam": mu.reshape(1, 36), "cov_prisparam": cov})
        cases = [(2, 200, 304, 4, "noise"), (1, 192, 192, 0, "smooth"), (1, 300, 210, 4, "mixed")]
        for ci, (b, h, w, border, kind) in enumerate(cases):
            g = torch.Generator().manual_seed(100 + ci)
            if kind == "noise":
                x = torch.rand(b, 3, h, w, generator=g)
            else:
                base = torch.rand(b, 3, h // 8 + 1, w // 8 + 1, generator=g)
                x = torch.nn.functional.interpolate(base, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
                if kind == "mixed":
                    x = (x + 0.15 * torch.rand(b, 3, h, w, generator=g)).clamp(0, 1)
            x = (x * 255).round() / 255          # images come from u8 files
            score = iqa._niqe_torch(x.clone(), border, path)
            out[f"x{ci}"] = (x * 255).round().to(torch.uint8).numpy()   # u8: x == stored / 255 in fp32
            out[f"border{ci}"] = np.int64(border)
            out[f"feat{ci}"] = captured["distparam"].numpy()
            out[f"niqe{ci}"] = np.atleast_1d(score.numpy())
            print(ci, kind, tuple(x.shape), "niqe", np.atleast_1d(score.numpy()), "features", tuple(captured["distparam"].shape))
    iqa._nanmean_torch = orig
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "niqe.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
