"""TEST INFRASTRUCTURE ONLY. numpy restatement of the reference's NIQE on tensors
(/root/reference/image_quality_assessment.py:886-998 `_niqe_torch` and the helpers it calls: `_fit_mscn_ipac_torch`
:861-... , `_get_mscn_feature_torch`, `_estimate_aggd_parameters_torch`, `_image_filter` with replicate padding,
`_image_resize_torch(scale=0.5, antialiasing=True)` = MATLAB's antialiased bicubic imresize, `_nanmean_torch`,
`_nancov_torch`; Y channel: imgproc.py:1815-1840 `rgb2ycbcr_torch`). Pinned by tests/golden/niqe.npz
(oracle/make_golden_niqe.py runs the unmodified reference with synthetic pristine statistics: the pretrained .mat is a
download that is not in the reference tree)."""
import numpy as np
from scipy.special import gammaln

_GAM = np.arange(0.2, 10 + 0.001, 0.001)                         # image_quality_assessment.py:808
_R_GAM = np.exp(2 * gammaln(2.0 / _GAM) - (gammaln(1.0 / _GAM) + gammaln(3.0 / _GAM)))


def y_channel_255(x_rgb: np.ndarray) -> np.ndarray:
    """[b, 3, h, w] fp32 in [0, 1] -> rounded Y in [16, 235] as float64 (imgproc.py:1829-1838, then :983-988)."""
    x = x_rgb.astype(np.float32)
    w = np.float32([65.481, 128.553, 24.966])
    y = (x[:, 0] * w[0] + x[:, 1] * w[1] + x[:, 2] * w[2] + np.float32(16.0)) / np.float32(255.0)
    return np.round(y * np.float32(255.0)).astype(np.float64)


def gaussian_window(size=7, sigma=7.0 / 6):
    m = (size - 1) / 2.0
    yy, xx = np.ogrid[-m:m + 1, -m:m + 1]
    h = np.exp(-(xx * xx + yy * yy) / (2.0 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    h /= h.sum()
    return h.astype(np.float32).astype(np.float64)                # the reference keeps the window as a float32 tensor


def filter_replicate(img: np.ndarray, win: np.ndarray) -> np.ndarray:
    """[b, h, w] correlation with replicate padding (ExactPadding2d(mode="replicate") + F.conv2d)."""
    r = win.shape[0] // 2
    p = np.pad(img, ((0, 0), (r, r), (r, r)), mode="edge")
    out = np.zeros_like(img)
    h, w = img.shape[1:]
    for dy in range(win.shape[0]):
        for dx in range(win.shape[1]):
            out += win[dy, dx] * p[:, dy:dy + h, dx:dx + w]
    return out


def aggd(block: np.ndarray):
    """block [n, bh, bw] -> (alpha, left_beta, right_beta), image_quality_assessment.py:790-836 with get_sigma=True."""
    left, right = block < 0, block > 0
    cl, cr = left.sum((-1, -2)).astype(np.float32).astype(np.float64), right.sum((-1, -2)).astype(np.float32).astype(np.float64)
    lstd = np.sqrt(((block * left) ** 2).sum((-1, -2)) / (cl + 1e-8))
    rstd = np.sqrt(((block * right) ** 2).sum((-1, -2)) / (cr + 1e-8))
    gamma_hat = lstd / rstd
    rhat = np.abs(block).mean((-1, -2)) ** 2 / (block ** 2).mean((-1, -2))
    rhat_norm = rhat * (gamma_hat ** 3 + 1) * (gamma_hat + 1) / (gamma_hat ** 2 + 1) ** 2
    pos = np.abs(_R_GAM[None, :] - rhat_norm[:, None]).argmin(-1)
    alpha = _GAM[pos]
    scale = np.sqrt(np.exp(gammaln(1 / alpha) - gammaln(3 / alpha)))
    return alpha, lstd * scale, rstd * scale


def mscn_features(blocks: np.ndarray) -> np.ndarray:
    """blocks [n, bh, bw] -> [n, 18] (image_quality_assessment.py:839-858)."""
    alpha, lb, rb = aggd(blocks)
    feats = [alpha, (lb + rb) / 2]
    for sh in ([0, 1], [1, 0], [1, 1], [1, -1]):
        shifted = np.roll(blocks, sh, axis=(1, 2))
        alpha, lb, rb = aggd(blocks * shifted)
        mean = (rb - lb) * np.exp(gammaln(2 / alpha) - gammaln(1 / alpha))
        feats.extend((alpha, mean, lb, rb))
    return np.stack(feats, -1)


def _cubic(x, a=-0.5):
    ax = np.abs(x)
    ax2, ax3 = ax * ax, ax * ax * ax
    c01 = ((a + 2) * ax3 - (a + 3) * ax2 + 1) * (ax <= 1)
    c12 = ((a * ax3) - (5 * a * ax2) + (8 * a * ax) - (4 * a)) * ((ax > 1) & (ax <= 2))
    return c01 + c12


def resize_half(img: np.ndarray) -> np.ndarray:
    """MATLAB imresize(img, 0.5) (bicubic with antialiasing) on [b, h, w], h and w even: ten taps with one weight set for
    every output (pos = 2 i + 0.5, base = 2 i - 4, dist = 4.5), symmetric padding of 4 (edge element repeated), height
    first, then width (image_quality_assessment.py:517-585)."""
    k = 10
    wgt = _cubic((4.5 - np.arange(k)) * 0.5)
    wgt = wgt / wgt.sum()

    def along(t, axis):
        t = np.moveaxis(t, axis, -1)
        n = t.shape[-1]
        p = np.concatenate([t[..., 3::-1], t, t[..., :-5:-1]], -1)       # [a,a,b,...]: boundary element used twice
        out = np.zeros(t.shape[:-1] + (n // 2,), t.dtype)
        for i in range(k):
            out += wgt[i] * p[..., i:i + n:2][..., :n // 2]
        return np.moveaxis(out, -1, axis)

    return along(along(img, 1), 2)


def features(y: np.ndarray, bs: int = 96) -> np.ndarray:
    """y [b, h, w] float64 (rounded Y * 255) -> [b, blocks, 36]."""
    b, h, w = y.shape
    nh, nw = h // bs, w // bs
    y = y[:, :nh * bs, :nw * bs]
    win = gaussian_window()
    out = []
    for scale in (1, 2):
        mu = filter_replicate(y, win)
        sd = np.sqrt(np.abs(filter_replicate(y * y, win) - mu * mu) + 1e-8)
        structdis = (y - mu) / (sd + 1)
        s = bs // scale
        blk = structdis.reshape(b, nh, s, nw, s).transpose(0, 1, 3, 2, 4).reshape(b * nh * nw, s, s)
        out.append(mscn_features(blk).reshape(b, nh * nw, 18))
        if scale == 1:
            y = resize_half(y / 255.0) * 255.0
    return np.concatenate(out, -1)


def niqe(x_rgb: np.ndarray, crop_border: int, mu_pris: np.ndarray, cov_pris: np.ndarray, bs: int = 96):
    """[b, 3, h, w] fp32 -> (scores [b], features [b, blocks, 36]); image_quality_assessment.py:886-998, 861-884."""
    if crop_border > 0:
        x_rgb = x_rgb[:, :, crop_border:-crop_border, crop_border:-crop_border]
    feat = features(y_channel_255(x_rgb), bs)
    scores = []
    for f in feat:
        ok = ~np.isnan(f).any(1)
        mu = np.nanmean(f, 0)
        cov = np.cov(f[ok], rowvar=False)
        inv = np.linalg.pinv((cov_pris + cov) / 2)
        d = mu_pris - mu
        scores.append(np.sqrt(d @ inv @ d))
    return np.array(scores), feat
