"""TEST INFRASTRUCTURE ONLY. numpy restatement of the reference dataset's training augmentation (dataset.py:66-79):
random_rotate (imgproc.py:1937-1963), random_horizontally_flip / random_vertically_flip (imgproc.py:1966-2001),
cv2.cvtColor(BGR2RGB) and image_to_tensor (imgproc.py:1540-1567), on an already decoded u8 image.

cv2.warpAffine with getRotationMatrix2D((w // 2, h // 2), 0 / 90 / 180 / 270, 1.0) is an exact pixel copy (its fixed-point
source coordinates snap to integers): dst(x, y) = src(M^-1 (x, y)), zeros outside the canvas. Pinned against the
reference's own functions (cv2 4.13.0) by tests/golden/augment.npz (oracle/make_golden_augment.py)."""
import numpy as np

ANGLES = (0, 90, 180, 270)


def augment(image_u8_bgr: np.ndarray, angle_index: int, hflip: bool, vflip: bool) -> np.ndarray:
    """[h, w, 3] u8 BGR -> [3, h, w] fp32 RGB in [0, 1]."""
    img = image_u8_bgr.astype(np.float32) / np.float32(255.)   # dataset.py:67
    h, w = img.shape[:2]
    cx, cy = w // 2, h // 2                                       # imgproc.py:1955-1956
    ys, xs = np.mgrid[0:h, 0:w]
    if hflip:                                                     # cv2.flip(image, 1), applied after the rotation
        xs = w - 1 - xs
    if vflip:                                                     # cv2.flip(image, 0)
        ys = h - 1 - ys
    if angle_index == 0:
        sx, sy = xs, ys
    elif angle_index == 1:
        sx, sy = cx + cy - ys, xs - cx + cy
    elif angle_index == 2:
        sx, sy = 2 * cx - xs, 2 * cy - ys
    else:
        sx, sy = ys + cx - cy, cx + cy - xs
    ok = (sx >= 0) & (sx < w) & (sy >= 0) & (sy < h)
    out = np.zeros_like(img)
    out[ok] = img[sy[ok], sx[ok]]
    return np.ascontiguousarray(out[:, :, ::-1].transpose(2, 0, 1))   # BGR -> RGB, HWC -> CHW


def pack_op(angle_index: int, hflip: bool, vflip: bool) -> int:
    return int(angle_index) | (int(bool(hflip)) << 2) | (int(bool(vflip)) << 3)
