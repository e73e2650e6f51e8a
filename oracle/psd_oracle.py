"""ORACLE (test infrastructure, not product): the Welch PSD the reference's frequency-domain analysis computes.

The reference calls ``scipy.signal.welch(image.flatten(), nperseg=256)`` (evaluate_Unet_diffusion/plot.py:155-157,
233-235, 287-290) and compares the spectra above half the maximum frequency (plot.py:159-165).  scipy is installed in
this image (and on the GPU box), so this part of the oracle IS the third-party implementation the reference uses —
parity pinned.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import welch


def welch_flat(image: np.ndarray):
    """(f, Pxx) of one image exactly as plot.py:155 computes it."""
    return welch(np.asarray(image).flatten(), nperseg=256)


def high_frequency_psd_mae(gt: np.ndarray, pred: np.ndarray, high_freq_threshold: float = 0.5) -> float:
    f, p_gt = welch_flat(gt)
    _, p_pr = welch_flat(pred)
    idx = f >= high_freq_threshold * np.max(f)
    return float(np.mean(np.abs(p_gt[idx] - p_pr[idx])))
