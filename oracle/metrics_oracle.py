"""ORACLE (test infrastructure, not product): PSNR / SSIM as the reference's evaluation scripts compute them.

PSNR follows the reference's own numpy code (evaluate_Unet_diffusion/evaluate_model.py:36-41) and the
skimage ``peak_signal_noise_ratio`` call (evaluate_SIDD/evaluate_SIDD.py:63).

SSIM — PARITY UNPINNED.  The arithmetic lives in scikit-image==0.22.0 (requirements.txt:98), which is not
vendored in the reference, not installed in this image, and pinned by no reference test or golden vector.
``structural_similarity`` below restates the published skimage 0.22 algorithm for the defaults the
reference's call sites use (evaluate_model.py:30-34: data_range=1.0, channel_axis=0;
evaluate_SIDD.py:64: data_range=2, channel_axis=-1): 7x7 uniform window (scipy.ndimage.uniform_filter,
mode='reflect'), K1=0.01, K2=0.03, sample covariance (cov_norm = 49/48), float32 arithmetic for float32
input, 3-pixel border crop, float64 mean per channel, channel means stored as float32 and averaged.
It is cross-checked in tests against an independent brute-force (explicit window loops, float64)
implementation of the same definition, and against analytic identities (SSIM(x,x)=1, symmetry).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def calculate_psnr(X: np.ndarray, Y: np.ndarray, data_range: float = 1.0) -> float:
    mse = np.mean((X - Y) ** 2)
    if mse == 0:
        return float("inf")
    return float(10 * np.log10((data_range ** 2) / mse))


def peak_signal_noise_ratio(image_true: np.ndarray, image_test: np.ndarray, data_range: float) -> float:
    # skimage: mse in float64 (images converted with _as_floats), 10*log10(R^2/mse)
    a = image_true.astype(np.float64)
    b = image_test.astype(np.float64)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    return float(10 * np.log10((data_range ** 2) / err))


def _ssim_plane(im1: np.ndarray, im2: np.ndarray, data_range: float, win_size: int = 7) -> float:
    K1, K2 = 0.01, 0.03
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    ftype = np.float32 if im1.dtype in (np.float16, np.float32) else np.float64
    im1 = im1.astype(ftype, copy=False)
    im2 = im2.astype(ftype, copy=False)
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    filt = dict(size=win_size)
    ux = uniform_filter(im1, **filt)
    uy = uniform_filter(im2, **filt)
    uxx = uniform_filter(im1 * im1, **filt)
    uyy = uniform_filter(im2 * im2, **filt)
    uxy = uniform_filter(im1 * im2, **filt)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    R = data_range
    C1 = (K1 * R) ** 2
    C2 = (K2 * R) ** 2
    A1, A2, B1, B2 = (2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2)
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    crop = S[pad:S.shape[0] - pad, pad:S.shape[1] - pad]
    return crop.mean(dtype=np.float64)


def structural_similarity(im1: np.ndarray, im2: np.ndarray, data_range: float, channel_axis=None) -> float:
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if channel_axis is None:
        return float(_ssim_plane(im1, im2, data_range))
    a = np.moveaxis(im1, channel_axis, 0)
    b = np.moveaxis(im2, channel_axis, 0)
    ftype = np.float32 if im1.dtype in (np.float16, np.float32) else np.float64
    per = np.empty(a.shape[0], dtype=ftype)
    for c in range(a.shape[0]):
        per[c] = _ssim_plane(a[c], b[c], data_range)
    return float(per.mean())


def ssim_bruteforce(im1: np.ndarray, im2: np.ndarray, data_range: float) -> float:
    """Independent float64 evaluation of the same definition with explicit 7x7 windows (small planes only)."""
    x = im1.astype(np.float64)
    y = im2.astype(np.float64)
    H, W = x.shape
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    acc, n = 0.0, 0
    for i in range(3, H - 3):
        for j in range(3, W - 3):
            wx = x[i - 3:i + 4, j - 3:j + 4]
            wy = y[i - 3:i + 4, j - 3:j + 4]
            mx, my = wx.mean(), wy.mean()
            vx = ((wx - mx) ** 2).sum() / 48.0
            vy = ((wy - my) ** 2).sum() / 48.0
            vxy = ((wx - mx) * (wy - my)).sum() / 48.0
            acc += ((2 * mx * my + C1) * (2 * vxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2))
            n += 1
    return acc / n
