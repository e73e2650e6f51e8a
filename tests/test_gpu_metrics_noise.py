"""GPU parity of the HBM-bound kernels: PSNR/SSIM reductions, Philox noise synthesis (bit-exact against the
CPU oracle at sigma = 10..50), the u8 <-> normalised boundary and the fused sampler step (bit-exact fp32)."""
import numpy as np
import pytest
import torch

import vub_image_denoising_b200 as b2
from oracle import metrics_oracle as mo
from oracle import noise_oracle as no
from oracle import rdunet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _images(n, c, h, w, seed, noise=0.1):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(n, c, h, w, generator=g) * 2 - 1
    b = (a + torch.randn(n, c, h, w, generator=g) * noise).clamp(-1, 1)
    return a, b


@pytest.mark.parametrize("shape", [(4, 3, 256, 256), (3, 3, 40, 72), (2, 1, 7, 7), (1, 3, 33, 45)])
@pytest.mark.parametrize("data_range", [1.0, 2.0])
def test_psnr_ssim_match_oracle(shape, data_range, built_lib):
    a, b = _images(*shape, seed=1)
    psnr, ssim = b2.metrics.batch_metrics(a.to(DEV), b.to(DEV), data_range)
    for i in range(shape[0]):
        ra, rb = a[i].numpy(), b[i].numpy()
        want_p = mo.calculate_psnr(ra, rb, data_range)
        want_s = mo.structural_similarity(ra, rb, data_range=data_range, channel_axis=0)
        assert float(psnr[i]) == pytest.approx(want_p, abs=1e-4)        # dB
        assert float(ssim[i]) == pytest.approx(want_s, abs=1e-5)


def test_metric_sums_are_bit_reproducible(built_lib):
    """Two-pass fixed-order reductions: repeated calls give identical bits (the atomicAdd version did not), and a
    batch-64 call equals the per-image calls bit for bit."""
    a, b = _images(64, 3, 256, 256, seed=5)
    A, B = a.to(DEV), b.to(DEV)
    p0, s0 = b2.metrics.batch_metrics(A, B, 1.0)
    for _ in range(5):
        p1, s1 = b2.metrics.batch_metrics(A, B, 1.0)
        assert torch.equal(p0, p1) and torch.equal(s0, s1)
    sse_all = b2.metrics.batch_sse(A, B)
    ss_all = b2.metrics.batch_ssim_planes(A.view(-1, 256, 256), B.view(-1, 256, 256), 1.0)
    for i in (0, 31, 63):
        assert torch.equal(b2.metrics.batch_ssim_planes(A[i], B[i], 1.0), ss_all[3 * i:3 * i + 3])
    assert torch.isfinite(sse_all).all()


@pytest.mark.parametrize("shape", [(2, 3, 256, 256), (3, 1, 40, 52), (1, 3, 16, 16), (2, 3, 64, 67)])
def test_welch_psd_matches_scipy(shape, built_lib):
    """f4: the Welch PSD of flattened images (plot.py:155-157) against scipy.signal.welch itself."""
    from oracle import psd_oracle as po
    g = torch.Generator().manual_seed(11)
    x = torch.rand(*shape, generator=g) * 2 - 1
    x[0] = torch.nn.functional.avg_pool2d(x[0:1], 3, stride=1, padding=1)[0]      # one image-like (low-passed) case
    f, pxx = b2.metrics.welch(x.to(DEV).flatten(1))
    assert pxx.shape == (shape[0], 129) and pxx.dtype == torch.float32
    for i in range(shape[0]):
        rf, rp = po.welch_flat(x[i].numpy())
        assert np.array_equal(f, rf) and rp.dtype == np.float32
        got = pxx[i].cpu().numpy()
        assert np.max(np.abs(got - rp)) <= 2e-5 * np.max(rp), f"image {i}: PSD differs from scipy by {np.max(np.abs(got - rp)):.3e}"
        assert np.max(np.abs(got - rp) / np.maximum(rp, 1e-12)) <= 2e-3
    # a single flattened image (the reference's call shape), and bit-reproducibility
    f1, p1 = b2.metrics.welch(x[0].to(DEV).flatten())
    assert p1.shape == (129,) and torch.equal(p1, pxx[0])
    with pytest.raises(ValueError):
        b2.metrics.welch(torch.zeros(100, device=DEV))


def test_high_frequency_psd_mae_matches_plot_py(built_lib):
    from oracle import psd_oracle as po
    a, b = _images(3, 3, 128, 128, seed=9, noise=0.2)
    got = b2.metrics.high_frequency_psd_mae(a.to(DEV), b.to(DEV)).cpu().numpy()
    for i in range(3):
        want = po.high_frequency_psd_mae(a[i].numpy(), b[i].numpy())
        assert got[i] == pytest.approx(want, rel=1e-3)


def test_reference_shaped_metric_calls(built_lib):
    a, b = _images(1, 3, 64, 64, seed=2)
    A, B = a[0].to(DEV), b[0].to(DEV)
    # evaluate_model.py:50-51 (CHW, data_range=1 on [-1,1] tensors)
    assert b2.metrics.calculate_psnr(A, B, 1.0) == pytest.approx(mo.calculate_psnr(a[0].numpy(), b[0].numpy(), 1.0), abs=1e-4)
    assert b2.metrics.calculate_ssim(A, B, 1.0, use_rgb=True) == pytest.approx(
        mo.structural_similarity(a[0].numpy(), b[0].numpy(), data_range=1.0, channel_axis=0), abs=1e-5)
    # evaluate_SIDD.py:59-64 (HWC numpy arrays, data_range=2)
    hwc_a, hwc_b = a[0].numpy().transpose(1, 2, 0), b[0].numpy().transpose(1, 2, 0)
    assert b2.metrics.peak_signal_noise_ratio(hwc_a, hwc_b, data_range=2) == pytest.approx(
        mo.peak_signal_noise_ratio(hwc_a, hwc_b, data_range=2), abs=1e-4)
    assert b2.metrics.structural_similarity(hwc_a, hwc_b, data_range=2, multichannel=True, channel_axis=-1) == \
        pytest.approx(mo.structural_similarity(hwc_a, hwc_b, data_range=2, channel_axis=-1), abs=1e-5)
    assert b2.metrics.calculate_psnr(A, A) == float("inf")
    assert b2.metrics.calculate_ssim(A, A, 1.0, use_rgb=True) == pytest.approx(1.0, abs=1e-6)
    with pytest.raises(ValueError):
        b2.metrics.calculate_ssim(A[:, :5], B[:, :5], 1.0, use_rgb=True)


def test_philox_normals_bit_exact(built_lib):
    for seed, stream, n in ((1234, 0, 1_000_003), (2 ** 40 + 17, 5, 4096), (0, 0, 5)):
        z = b2.noise.philox_normal(n, seed, stream, device=DEV).cpu().numpy()
        assert np.array_equal(z.view(np.uint32), no.normals(n, seed, stream).view(np.uint32))


@pytest.mark.parametrize("sigma", [10, 20, 30, 40, 50])
def test_noise_synthesis_bit_exact(sigma, built_lib):
    rng = np.random.default_rng(sigma)
    clean = rng.integers(0, 256, size=(3, 64, 96, 3), dtype=np.uint8)
    n_u8, n_norm, c_norm = b2.noise.add_gaussian_noise(torch.from_numpy(clean).to(DEV), float(sigma), seed=777)
    r_u8, r_norm, r_cnorm = no.degrade(clean, float(sigma), seed=777)
    assert np.array_equal(n_u8.cpu().numpy(), r_u8)
    assert np.array_equal(n_norm.cpu().numpy().view(np.uint32), r_norm.view(np.uint32))
    assert np.array_equal(c_norm.cpu().numpy().view(np.uint32), r_cnorm.view(np.uint32))


def test_noise_per_image_sigma_cycle_and_gray(built_lib):
    rng = np.random.default_rng(0)
    clean = rng.integers(0, 256, size=(5, 32, 32, 3), dtype=np.uint8)
    sig = [10.0, 20.0, 30.0, 40.0, 50.0]                  # evaluate_model.py:315-318 noise levels
    n_u8, n_norm, _ = b2.noise.add_gaussian_noise(torch.from_numpy(clean).to(DEV), sig, seed=5, stream_id=2)
    r_u8, r_norm, _ = no.degrade(clean, np.array(sig, dtype=np.float32), seed=5, stream_id=2)
    assert np.array_equal(n_u8.cpu().numpy(), r_u8) and np.array_equal(n_norm.cpu().numpy(), r_norm)
    gray = rng.integers(0, 256, size=(2, 16, 16, 1), dtype=np.uint8)
    g_u8, g_norm, _ = b2.noise.add_gaussian_noise(torch.from_numpy(gray).to(DEV), 25.0, seed=9)
    r_u8, r_norm, _ = no.degrade(gray, 25.0, seed=9)
    assert np.array_equal(g_u8.cpu().numpy(), r_u8) and np.array_equal(g_norm.cpu().numpy(), r_norm)


def test_u8_boundary_round_trip(built_lib):
    rng = np.random.default_rng(1)
    u8 = rng.integers(0, 256, size=(2, 24, 40, 3), dtype=np.uint8)
    norm = b2.noise.u8_to_normalized(torch.from_numpy(u8).to(DEV))
    chw = u8.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)
    assert np.array_equal(norm.cpu().numpy(), (chw - np.float32(0.5)) / np.float32(0.5))
    back = b2.noise.normalized_to_u8(norm)
    assert np.array_equal(back.cpu().numpy(), no.norm_to_u8(norm.cpu().numpy()))
    # quantise(normalise(u8)) is the identity up to the truncation of values that land just below an integer
    assert np.abs(back.cpu().numpy().astype(int) - u8.astype(int)).max() <= 1
    wild = torch.tensor([-1.5, -1.0, 0.0, 0.999, 1.0, 3.0], device=DEV).view(1, 1, 1, 6)
    assert np.array_equal(b2.noise.normalized_to_u8(wild).cpu().numpy(), no.norm_to_u8(wild.cpu().numpy()))


@pytest.mark.parametrize("n_shape", [(2, 3, 16, 16), (1, 3, 5, 7), (16, 3, 256, 256)])
def test_sampler_step_bit_exact(n_shape, built_lib):
    g = torch.Generator().manual_seed(0)
    x, u1, u2, y = (torch.randn(*n_shape, generator=g) for _ in range(4))
    for t, T in ((20, 20), (7, 20), (1, 20), (3, 4)):
        a_t, a_p = t / T, (t - 1) / T
        f32 = lambda v: float(np.float32(v))  # noqa: E731
        got = torch.ops.b200dn.sampler_step(x.to(DEV), u1.to(DEV), u2.to(DEV), y.to(DEV),
                                            f32(1 - a_t), f32(a_t), f32(1 - a_p), f32(a_p))
        want = orc.sampler_step(x, u1, u2, y, t, T)
        assert torch.equal(got.cpu(), want)
