"""Restatement of the three scikit-image metrics the reference calls (src/util/error.py:10-12, :23-84) and of its
data-range rule.  Test infrastructure only (see ``oracle/__init__.py``).

PARITY UNPINNED: scikit-image is a third-party dependency of the reference (``requirements.txt:9``, no version pin)
that is not installed in this image and is not vendored under /root/reference, and the reference has no tests or
golden values for these calls.  The functions below follow the published algorithms of
``skimage.metrics.{peak_signal_noise_ratio, structural_similarity, normalized_root_mse}`` with the defaults the
reference uses (2-D float images, ``data_range`` given, ``win_size=7``, uniform filter, ``K1=0.01``, ``K2=0.03``,
sample covariance, border of ``(win_size-1)//2`` pixels cropped from the SSIM mean; NRMSE with euclidean
normalisation).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def data_range(original: np.ndarray, predicted: np.ndarray) -> float:
    """src/util/error.py:23-38: max over both images minus min over both images."""
    return float(max(original.max(), predicted.max()) - min(original.min(), predicted.min()))


def psnr(original: np.ndarray, predicted: np.ndarray) -> float:
    """error.py:41-54 -> skimage.metrics.peak_signal_noise_ratio(..., data_range=...)."""
    o, p = original.astype(np.float64), predicted.astype(np.float64)
    mse = np.mean((o - p) ** 2)
    return float(10.0 * np.log10(data_range(original, predicted) ** 2 / mse))


def nrmse(original: np.ndarray, predicted: np.ndarray) -> float:
    """error.py:73-84 -> skimage.metrics.normalized_root_mse (euclidean normalisation)."""
    o, p = original.astype(np.float64), predicted.astype(np.float64)
    return float(np.sqrt(np.mean((o - p) ** 2)) / np.sqrt(np.mean(o ** 2)))


def ssim(original: np.ndarray, predicted: np.ndarray, win_size: int = 7, k1: float = 0.01, k2: float = 0.03) -> float:
    """error.py:57-70 -> skimage.metrics.structural_similarity(..., data_range=...) with skimage's defaults."""
    x, y = original.astype(np.float64), predicted.astype(np.float64)
    r = data_range(original, predicted)
    npix = win_size ** 2
    cov_norm = npix / (npix - 1.0)          # sample covariance
    ux, uy = uniform_filter(x, win_size), uniform_filter(y, win_size)
    uxx, uyy, uxy = uniform_filter(x * x, win_size), uniform_filter(y * y, win_size), uniform_filter(x * y, win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * r) ** 2, (k2 * r) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win_size - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())


def ssim_direct(original: np.ndarray, predicted: np.ndarray, win_size: int = 7, k1: float = 0.01, k2: float = 0.03) -> float:
    """A second, independent formulation of the same SSIM: every win x win window is materialised
    (``sliding_window_view``) and its sample mean / variance / covariance are computed directly from centred values in
    fp64 -- no uniform filter, no ``E[x^2] - E[x]^2`` difference, no border handling (only windows fully inside the
    image exist, which is exactly the set skimage keeps after cropping ``(win_size-1)//2`` pixels).  Used to pin
    ``ssim`` above and the CUDA kernel against something that shares no code path with either."""
    from numpy.lib.stride_tricks import sliding_window_view

    x, y = original.astype(np.float64), predicted.astype(np.float64)
    r = data_range(original, predicted)
    wx = sliding_window_view(x, (win_size, win_size)).reshape(x.shape[0] - win_size + 1, x.shape[1] - win_size + 1, -1)
    wy = sliding_window_view(y, (win_size, win_size)).reshape(wx.shape)
    n = win_size * win_size
    mx, my = wx.mean(axis=-1), wy.mean(axis=-1)
    dx, dy = wx - mx[..., None], wy - my[..., None]
    vx, vy, vxy = (dx * dx).sum(-1) / (n - 1), (dy * dy).sum(-1) / (n - 1), (dx * dy).sum(-1) / (n - 1)
    c1, c2 = (k1 * r) ** 2, (k2 * r) ** 2
    s = ((2 * mx * my + c1) * (2 * vxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
    return float(s.mean())
