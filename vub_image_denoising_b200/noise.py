"""Device-side noisy-patch synthesis and the u8 <-> normalised-fp32 data-format boundary.

Reference arithmetic (host numpy / torchvision in the reference):
  * ``noisy = clip(float32(patch) + N(0, sigma), 0, 255).astype(uint8)``   dataset_creation/custom_dataset.py:83-87
  * ``ToTensor`` (HWC u8 -> CHW fp32 /255) + ``Normalize(0.5, 0.5)``         dataset_creation/data_loader.py:35-38,
                                                                            evaluate_SIDD/evaluate_SIDD.py:23-26
  * ``(y + 1) / 2 -> clip(. * 255, 0, 255).astype(uint8)``                  evaluate_SIDD/benchmark.py:42-44

The reference draws from numpy's unseeded global MT19937, so there is no bit stream to reproduce; the
generator here is Philox4x32-10 + Box-Muller as specified in csrc/philox_normal.h, reproducible bit for
bit on the device and in the CPU oracle for a given (seed, stream_id).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch

from . import _lib

__all__ = ["add_gaussian_noise", "philox_normal", "u8_to_normalized", "normalized_to_u8"]


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA (sm_100) tensor required; there is no CPU fallback")


def add_gaussian_noise(clean_u8: torch.Tensor, sigma: Union[float, Sequence[float], torch.Tensor], seed: int,
                       stream_id: int = 0, return_u8: bool = True, return_clean: bool = True
                       ) -> Tuple[Optional[torch.Tensor], torch.Tensor, Optional[torch.Tensor]]:
    """clean_u8: uint8 [B, H, W, C] (C = 1 or 3, HWC as PIL / the SIDD blocks store it).

    Returns ``(noisy_u8 [B,H,W,C] | None, noisy_norm fp32 [B,C,H,W] in [-1,1], clean_norm | None)``.
    ``sigma`` is one value or one per image (the reference cycles sigma over the dataset index,
    custom_dataset.py:68-83)."""
    _need_cuda(clean_u8, "add_gaussian_noise")
    if clean_u8.dtype != torch.uint8 or clean_u8.dim() != 4:
        raise RuntimeError("clean_u8 must be a uint8 [B, H, W, C] tensor")
    clean_u8 = clean_u8.contiguous()
    B, H, W, Cn = clean_u8.shape
    dev = clean_u8.device
    if isinstance(sigma, torch.Tensor):
        sig = sigma.to(device=dev, dtype=torch.float32).reshape(-1)
    elif isinstance(sigma, (int, float)):
        sig = torch.full((B,), float(sigma), dtype=torch.float32, device=dev)
    else:
        sig = torch.tensor([float(s) for s in sigma], dtype=torch.float32, device=dev)
    if sig.numel() == 1 and B > 1:
        sig = sig.expand(B)
    if sig.numel() != B:
        raise RuntimeError(f"sigma must have 1 or B={B} entries, got {sig.numel()}")
    sig = sig.contiguous()
    noisy_u8 = torch.empty_like(clean_u8) if return_u8 else None
    noisy = torch.empty((B, Cn, H, W), dtype=torch.float32, device=dev)
    clean = torch.empty_like(noisy) if return_clean else None
    with torch.cuda.device(dev):
        rc = _lib.lib().b200dn_gauss_noise_u8(
            clean_u8.data_ptr(), B, H, W, Cn, sig.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF,
            noisy_u8.data_ptr() if return_u8 else None, noisy.data_ptr(), clean.data_ptr() if return_clean else None,
            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "gauss_noise_u8")
    return noisy_u8, noisy, clean


def philox_normal(n: int, seed: int, stream_id: int = 0, device=None) -> torch.Tensor:
    """n standard normals (fp32) from the repo's generator spec — used to pin the RNG against the oracle."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("philox_normal: CUDA (sm_100) device required; there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    z = torch.empty(n, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = _lib.lib().b200dn_philox_normal(z.data_ptr(), n, int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF,
                                             torch.cuda.current_stream(device).cuda_stream)
    _lib.check(rc, "philox_normal")
    return z


def u8_to_normalized(img_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [B, H, W, C] -> fp32 [B, C, H, W] = (x/255 - 0.5)/0.5."""
    _need_cuda(img_u8, "u8_to_normalized")
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4:
        raise RuntimeError("expected a uint8 [B, H, W, C] tensor")
    img_u8 = img_u8.contiguous()
    B, H, W, Cn = img_u8.shape
    out = torch.empty((B, Cn, H, W), dtype=torch.float32, device=img_u8.device)
    with torch.cuda.device(img_u8.device):
        rc = _lib.lib().b200dn_u8_to_norm(img_u8.data_ptr(), B, H, W, Cn, out.data_ptr(),
                                          torch.cuda.current_stream(img_u8.device).cuda_stream)
    _lib.check(rc, "u8_to_norm")
    return out


def normalized_to_u8(img: torch.Tensor) -> torch.Tensor:
    """fp32 [B, C, H, W] in [-1, 1] -> uint8 [B, H, W, C] with the benchmark.py:42-44 arithmetic."""
    _need_cuda(img, "normalized_to_u8")
    if img.dim() != 4:
        raise RuntimeError("expected a [B, C, H, W] tensor")
    img = img.detach().to(torch.float32).contiguous()
    B, Cn, H, W = img.shape
    out = torch.empty((B, H, W, Cn), dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        rc = _lib.lib().b200dn_norm_to_u8(img.data_ptr(), B, H, W, Cn, out.data_ptr(),
                                          torch.cuda.current_stream(img.device).cuda_stream)
    _lib.check(rc, "norm_to_u8")
    return out
