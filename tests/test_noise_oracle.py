"""CPU: the noise-synthesis oracle. Philox4x32-10 is pinned by the Random123 known-answer vectors; the
degradation arithmetic is pinned against a numpy restatement of the reference's own lines
(dataset_creation/custom_dataset.py:84-86, dataset_creation/data_loader.py:35-38)."""
import numpy as np
import pytest

from oracle import noise_oracle as no

KAT = [  # Random123 kat_vectors, philox4x32-10: counter, key, expected
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    assert [int(v) for v in no.philox4x32_10(ctr, key)] == want


def test_normals_are_standard_and_reproducible():
    z = no.normals(1_000_000, seed=1234)
    assert np.array_equal(z, no.normals(1_000_000, seed=1234))
    assert not np.array_equal(z, no.normals(1_000_000, seed=1235))
    assert not np.array_equal(z[:1000], no.normals(1000, seed=1234, stream_id=1))
    assert np.all(np.isfinite(z))
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1) < 3e-3
    assert abs((z ** 3).mean()) < 1e-2 and abs((z ** 4).mean() - 3) < 3e-2
    # prefix property: element i does not depend on n
    assert np.array_equal(no.normals(10, 7), no.normals(1001, 7)[:10])
    # Box-Muller pairs are uncorrelated
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5e-3


def test_log_sincos_accuracy():
    """The fixed polynomials are accurate to a few ulp: the radius/angle identities hold."""
    z = no.normals(400_000, seed=99).astype(np.float64)
    r2 = z[0::2] ** 2 + z[1::2] ** 2          # = -2 ln u1 -> exponential(1/2)
    assert abs(r2.mean() - 2.0) < 2e-2
    ang = np.arctan2(z[1::2], z[0::2])
    assert abs(ang.mean()) < 2e-2 and abs(ang.std() - np.pi / np.sqrt(3)) < 1e-2


@pytest.mark.parametrize("sigma", [10, 20, 30, 40, 50])
def test_degrade_matches_reference_arithmetic(sigma):
    rng = np.random.default_rng(0)
    clean = rng.integers(0, 256, size=(2, 16, 24, 3), dtype=np.uint8)
    noisy_u8, noisy, clean_n = no.degrade(clean, float(sigma), seed=42)
    z = no.normals(clean.size, 42).reshape(clean.shape)
    # custom_dataset.py:84-86 with our fp32 noise in place of np.random.normal
    ref = np.array(clean, dtype=np.float32)
    ref += (np.float32(sigma) * z).astype(np.float32)
    ref_u8 = np.clip(ref, 0, 255).astype(np.uint8)
    assert np.array_equal(noisy_u8, ref_u8)
    # ToTensor + Normalize (data_loader.py:35-38): HWC u8 -> CHW float32 /255, then (x - 0.5) / 0.5
    chw = ref_u8.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)
    assert np.array_equal(noisy, (chw - np.float32(0.5)) / np.float32(0.5))
    chw_c = clean.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)
    assert np.array_equal(clean_n, (chw_c - np.float32(0.5)) / np.float32(0.5))
    # a sigma-dependent share of samples saturates; none leave [0,255]
    assert noisy_u8.min() >= 0 and noisy_u8.max() <= 255
    assert (noisy_u8 != clean).mean() > 0.8


def test_per_image_sigma_and_quantiser():
    clean = np.full((3, 8, 8, 3), 128, dtype=np.uint8)
    n_u8, _, _ = no.degrade(clean, np.array([0.0, 10.0, 50.0], dtype=np.float32), seed=3)
    assert np.array_equal(n_u8[0], clean[0])
    assert n_u8[1].astype(int).std() < n_u8[2].astype(int).std()
    img = np.array([-1.0, -0.999, 0.0, 0.5, 1.0, 1.2, -1.3], dtype=np.float32).reshape(1, 1, 1, 7)
    ref = np.clip(((img + 1) / 2) * 255, 0, 255).astype(np.uint8).transpose(0, 2, 3, 1)   # benchmark.py:42-44
    assert np.array_equal(no.norm_to_u8(img), ref)
