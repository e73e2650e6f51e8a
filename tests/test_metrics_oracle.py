"""CPU: the PSNR/SSIM oracle. PSNR follows the reference's own numpy code; SSIM (skimage 0.22, not vendored:
parity unpinned) is cross-checked against an independent brute-force evaluation of the same definition."""
import numpy as np
import pytest

from oracle import metrics_oracle as mo


def _pair(shape, seed, noise=0.1):
    rng = np.random.default_rng(seed)
    a = rng.random(shape, dtype=np.float32) * 2 - 1
    b = np.clip(a + rng.normal(0, noise, shape).astype(np.float32), -1, 1)
    return a, b


def test_psnr_definition():
    a, b = _pair((3, 32, 32), 0)
    mse = float(np.mean((a.astype(np.float64) - b) ** 2))
    assert mo.calculate_psnr(a, b, 1.0) == pytest.approx(10 * np.log10(1 / mse), abs=1e-4)
    assert mo.peak_signal_noise_ratio(a, b, data_range=2) == pytest.approx(10 * np.log10(4 / mse), abs=1e-9)
    assert mo.calculate_psnr(a, a) == float("inf")


@pytest.mark.parametrize("data_range", [1.0, 2.0])
def test_ssim_matches_bruteforce(data_range):
    a, b = _pair((24, 20), 1)
    fast = mo.structural_similarity(a, b, data_range=data_range)
    slow = mo.ssim_bruteforce(a, b, data_range)
    assert fast == pytest.approx(slow, abs=2e-5)


def test_ssim_properties():
    a, b = _pair((3, 32, 40), 2)
    assert mo.structural_similarity(a, a, data_range=1.0, channel_axis=0) == pytest.approx(1.0, abs=1e-6)
    s_ab = mo.structural_similarity(a, b, data_range=1.0, channel_axis=0)
    s_ba = mo.structural_similarity(b, a, data_range=1.0, channel_axis=0)
    assert s_ab == pytest.approx(s_ba, abs=1e-6) and 0 < s_ab < 1
    # channel_axis=-1 on HWC equals channel_axis=0 on CHW (evaluate_SIDD.py:59-64 vs evaluate_model.py:30-34)
    hwc = mo.structural_similarity(a.transpose(1, 2, 0), b.transpose(1, 2, 0), data_range=1.0, channel_axis=-1)
    assert hwc == pytest.approx(s_ab, abs=1e-7)
    # more noise, less similarity
    _, c = _pair((3, 32, 40), 2, noise=0.4)
    assert mo.structural_similarity(a, c, data_range=1.0, channel_axis=0) < s_ab
    with pytest.raises(ValueError):
        mo.structural_similarity(a[:, :5], b[:, :5], data_range=1.0, channel_axis=0)
