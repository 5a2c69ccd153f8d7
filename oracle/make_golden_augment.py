"""Writes tests/golden/augment.npz from the UNMODIFIED reference functions (build container only):
    python -m oracle.make_golden_augment
For every image and every (angle, hflip, vflip) the chain of dataset.py:67-79 is run through the reference's own
imgproc.random_rotate / random_horizontally_flip / random_vertically_flip (probabilities forced to 0 / 1 so that every
branch is recorded), cv2.cvtColor and imgproc.image_to_tensor."""
import os

import numpy as np

from . import refshim


def main():
    import cv2
    _, rip, _ = refshim.load()
    rng = np.random.default_rng(7)
    out = {}
    for idx, (h, w) in enumerate([(8, 8), (9, 9), (6, 11), (11, 6), (40, 40), (37, 52)]):
        img_u8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out[f"img{idx}"] = img_u8
        image = img_u8.astype(np.float32) / 255.
        for ai, ang in enumerate([0, 90, 180, 270]):
            for hf in (0, 1):
                for vf in (0, 1):
                    t = rip.random_rotate(image, [ang])
                    t = rip.random_horizontally_flip(t, 1.0 if hf else 0.0)
                    t = rip.random_vertically_flip(t, 1.0 if vf else 0.0)
                    t = cv2.cvtColor(t, cv2.COLOR_BGR2RGB)
                    out[f"out{idx}_{ai}_{hf}_{vf}"] = rip.image_to_tensor(t, False, False).numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "augment.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
