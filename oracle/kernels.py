"""Oracle for blur-kernel synthesis: numpy float64 restatement. TEST INFRASTRUCTURE ONLY.

Follows /root/reference/imgproc.py: _mesh_grid :72-90, _calculate_rotate_sigma_matrix :170-186,
_calculate_probability_density :189-203, the three kernel families :225-327, generate_sinc_kernel :576-603
(scipy.special.j1, scipy 1.18.1). Pinned by tests/golden/kernels.npz.
"""
import numpy as np
from scipy import special


def _quad(k, sigma_x, sigma_y, theta, isotropic):
    ax = np.arange(-k // 2 + 1.0, k // 2 + 1.0)
    xx, yy = np.meshgrid(ax, ax)  # first quadratic-form coordinate = column offset
    if isotropic:
        sig = np.array([[sigma_x ** 2, 0], [0, sigma_x ** 2]])
    else:
        d = np.array([[sigma_x ** 2, 0], [0, sigma_y ** 2]])
        u = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
        sig = u @ d @ u.T
    inv = np.linalg.inv(sig)
    g = np.stack([xx, yy], -1)
    return np.sum((g @ inv) * g, 2)


def gaussian(k, sigma_x, sigma_y=None, theta=0.0, isotropic=True):
    v = np.exp(-0.5 * _quad(k, sigma_x, sigma_y, theta, isotropic))
    return v / v.sum()


def generalized(k, sigma_x, sigma_y, theta, beta, isotropic=True):
    v = np.exp(-0.5 * np.power(_quad(k, sigma_x, sigma_y, theta, isotropic), beta))
    return v / v.sum()


def plateau(k, sigma_x, sigma_y, theta, beta, isotropic=True):
    v = np.reciprocal(np.power(_quad(k, sigma_x, sigma_y, theta, isotropic), beta) + 1)
    return v / v.sum()


def sinc(cutoff, k, padding=0):
    c = (k - 1) / 2
    i, j = np.meshgrid(np.arange(k), np.arange(k), indexing="ij")
    r = np.sqrt((i - c) ** 2 + (j - c) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = cutoff * special.j1(cutoff * r) / (2 * np.pi * r)
    v[(k - 1) // 2, (k - 1) // 2] = cutoff ** 2 / (4 * np.pi)
    v = v / v.sum()
    if padding > k:
        p = (padding - k) // 2
        v = np.pad(v, ((p, p), (p, p)))
    return v


def from_params(p, pad=0):
    """Evaluate one parameter dict as produced by resr_b200.imgproc.draw_mixed_kernel_params."""
    t, k = p["type"], p["kernel_size"]
    if t == "sinc":
        v = sinc(p["cutoff"], k)
    elif t == "delta":
        v = np.zeros((k, k))
        v[k // 2, k // 2] = 1
    elif t == "gaussian":
        v = gaussian(k, p["sigma_x"], p["sigma_y"], p["theta"], p["isotropic"])
    elif t == "generalized":
        v = generalized(k, p["sigma_x"], p["sigma_y"], p["theta"], p["beta"], p["isotropic"])
    else:
        v = plateau(k, p["sigma_x"], p["sigma_y"], p["theta"], p["beta"], p["isotropic"])
    if pad > k:
        q = (pad - k) // 2
        v = np.pad(v, ((q, q), (q, q)))
    return v
