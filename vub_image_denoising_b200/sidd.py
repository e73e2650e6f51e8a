"""SIDD sRGB paths of the reference, batched on the device.

Mirrors, with the reference's own names and argument meaning:

* ``evaluate_SIDD/benchmark.py:32-46``   ``my_srgb_denoiser(x)``: uint8 HWC block -> ToTensor -> Normalize(0.5, 0.5)
  -> ``model.improved_sampling`` -> ``(y + 1) / 2`` -> ``clip(y * 255, 0, 255).astype(uint8)``;
* ``evaluate_SIDD/benchmark.py:48-58``   ``array_to_base64string`` / ``base64string_to_array``;
* ``evaluate_SIDD/benchmark.py:79-103``  the block loop over ``[40, 32, 256, 256, 3]`` and ``SubmitSrgb.csv``
  (columns ``ID``, ``BLOCK``; blocks in row-major (image, patch) order);
* ``evaluate_SIDD/evaluate_SIDD.py:18-41,43-75``  ``SIDDMatDataset`` normalisation and ``evaluate_model``:
  per-patch ``peak_signal_noise_ratio(gt, out, data_range=2)`` / ``structural_similarity(gt, out, data_range=2,
  channel_axis=-1)`` on [-1, 1] HWC arrays, then the mean over patches.

The reference runs one 256x256 block per model call (``batch_size=1``, ``evaluate_SIDD.py:116``) with a host round
trip per block.  Here blocks go through in batches: one H2D of the uint8 blocks, normalisation / sampler /
quantisation / metrics on the device, one D2H of the uint8 result.  A batched call returns bit-identical uint8 blocks
to a per-block loop (every op of the path is per-sample).  With ``torch.distributed`` initialised, ``evaluate_sidd``
shards the flat block list over the ranks and combines (sum PSNR, sum SSIM, count) with ONE all-reduce.
There is no CPU path: the model must live on a CUDA device.
"""
from __future__ import annotations

import base64
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import metrics, noise, sharding

__all__ = ["array_to_base64string", "base64string_to_array", "flatten_blocks", "denoise_blocks_srgb",
           "my_srgb_denoiser", "set_default_model", "submission_rows", "write_submission_csv", "evaluate_sidd"]

ArrayLike = Union[np.ndarray, torch.Tensor]


# ------------------------------------------------------------------ host encoders (benchmark.py:48-58)
def array_to_base64string(x: np.ndarray) -> str:
    """``base64(x.tobytes())`` as utf-8 text (benchmark.py:48-52)."""
    return base64.b64encode(np.ascontiguousarray(x).tobytes()).decode("utf-8")


def base64string_to_array(base64string: str, array_dtype, array_shape) -> np.ndarray:
    """Inverse of :func:`array_to_base64string` (benchmark.py:54-58)."""
    return np.frombuffer(base64.b64decode(base64string), dtype=array_dtype).reshape(array_shape)


# ------------------------------------------------------------------ block bookkeeping
def flatten_blocks(blocks: ArrayLike) -> Tuple[torch.Tensor, Tuple[int, ...]]:
    """uint8 blocks ``[H, W, C]``, ``[N, H, W, C]`` or ``[I, P, H, W, C]`` (the .mat layout) -> (flat ``[N, H, W, C]``
    CPU tensor in the reference's loop order ``for i: for j:``, original shape)."""
    t = torch.from_numpy(np.ascontiguousarray(blocks)) if isinstance(blocks, np.ndarray) else blocks
    if t.dtype != torch.uint8:
        raise RuntimeError(f"SIDD blocks must be uint8, got {t.dtype}")
    if t.dim() not in (3, 4, 5):
        raise RuntimeError(f"expected [H,W,C], [N,H,W,C] or [I,P,H,W,C] blocks, got shape {tuple(t.shape)}")
    shape = tuple(t.shape)
    H, W, C = shape[-3:]
    if C != 3:
        raise RuntimeError(f"sRGB blocks have 3 channels, got {C}")
    if H % 8 or W % 8:
        raise RuntimeError(f"block size {H}x{W} must be divisible by 8 (three 2x2 down/up stages)")
    return t.reshape(-1, H, W, C).contiguous(), shape


def _model_device(model: torch.nn.Module) -> torch.device:
    p = next(model.parameters(), None)
    if p is None or p.device.type != "cuda":
        raise RuntimeError("the SIDD path has no CPU fallback: move the model to a CUDA device")
    return p.device


def _denoise(model, x: torch.Tensor, batch: int) -> torch.Tensor:
    """DiffusionModel -> improved_sampling (benchmark.py:39, evaluate_SIDD.py:56); a bare RDUNet -> forward.
    A ragged last batch is padded to `batch` samples (every op is per-sample, so the kept rows are unchanged): the
    sampler's captured CUDA graph and workspaces are keyed on the batch size."""
    n = x.shape[0]
    if n < batch:
        x = torch.cat([x, x[:1].expand(batch - n, -1, -1, -1)], 0)
    y = model.improved_sampling(x) if hasattr(model, "improved_sampling") else model(x)
    return y[:n]


# ------------------------------------------------------------------ benchmark.py:32-46 / 79-92, batched
@torch.no_grad()
def denoise_blocks_srgb(model: torch.nn.Module, blocks: ArrayLike, batch: int = 32) -> np.ndarray:
    """Denoise uint8 sRGB blocks; returns a uint8 array of the input's shape (benchmark.py:79-92 without the
    per-block host round trip)."""
    if batch < 1:
        raise RuntimeError("batch must be >= 1")
    dev = _model_device(model)
    flat, shape = flatten_blocks(blocks)
    n = flat.shape[0]
    src = flat if flat.is_cuda else (flat if flat.is_pinned() else flat.pin_memory())
    out = torch.empty(flat.shape, dtype=torch.uint8).pin_memory()
    with torch.cuda.device(dev):
        for lo in range(0, n, batch):
            hi = min(n, lo + batch)
            x_u8 = src[lo:hi].to(dev, non_blocking=True)
            den = _denoise(model, noise.u8_to_normalized(x_u8), min(batch, n))   # ToTensor + Normalize, sampler
            out[lo:hi].copy_(noise.normalized_to_u8(den), non_blocking=True)   # (y+1)/2 -> clip(.*255) -> uint8
        torch.cuda.current_stream(dev).synchronize()
    return out.numpy().reshape(shape)


_default_model: Optional[torch.nn.Module] = None


def set_default_model(model: Optional[torch.nn.Module]) -> None:
    """The model ``my_srgb_denoiser(x)`` uses when called with one argument — benchmark.py keeps it in a module-level
    global (``model``, benchmark.py:23-27); ``shim.install(model=...)`` calls this."""
    global _default_model
    _default_model = model


def my_srgb_denoiser(x: np.ndarray, model: Optional[torch.nn.Module] = None) -> np.ndarray:
    """One uint8 ``[H, W, 3]`` block in, one out (benchmark.py:32-46).  ``my_srgb_denoiser(x)`` — the reference's
    one-argument form — uses the model registered with :func:`set_default_model`."""
    model = model if model is not None else _default_model
    if model is None:
        raise RuntimeError("my_srgb_denoiser(x): no model registered; call sidd.set_default_model(model) / "
                           "shim.install(model=model), or pass the model as second argument")
    if x.ndim != 3:
        raise RuntimeError(f"my_srgb_denoiser takes one [H,W,3] block, got shape {x.shape}")
    return denoise_blocks_srgb(model, x, batch=1)


# ------------------------------------------------------------------ SubmitSrgb.csv (benchmark.py:94-103)
def submission_rows(out_blocks: ArrayLike) -> List[Tuple[int, str]]:
    """(ID, BLOCK) rows: block k of the flattened (image, patch) order, base64 of its raw uint8 bytes."""
    flat, _ = flatten_blocks(out_blocks)
    arr = flat.cpu().numpy()
    return [(k, array_to_base64string(arr[k])) for k in range(arr.shape[0])]


def write_submission_csv(path: str, out_blocks: ArrayLike) -> int:
    """Write the Kaggle submission file the way ``DataFrame.to_csv(index=False)`` does for two plain columns."""
    rows = submission_rows(out_blocks)
    with open(path, "w", newline="") as f:
        f.write("ID,BLOCK\n")
        for k, s in rows:
            f.write(f"{k},{s}\n")
    return len(rows)


# ------------------------------------------------------------------ evaluate_SIDD.py:43-75, batched + sharded
@torch.no_grad()
def evaluate_sidd(model: torch.nn.Module, noisy_blocks: ArrayLike, gt_blocks: ArrayLike, batch: int = 32,
                  group=None, return_denoised: bool = False, presharded: bool = False) -> Dict[str, object]:
    """Average PSNR / SSIM (data_range = 2 on [-1, 1] images, SSIM over the channel-last RGB image) of the denoised
    noisy blocks against the ground-truth blocks.  Under ``torch.distributed`` every rank evaluates its contiguous
    shard of the flattened block list (or, with ``presharded``, exactly the blocks it was given) and the sums are
    combined with one all-reduce."""
    dev = _model_device(model)
    noisy, shape = flatten_blocks(noisy_blocks)
    gt, gshape = flatten_blocks(gt_blocks)
    if shape != gshape:
        raise RuntimeError(f"noisy blocks {shape} and ground-truth blocks {gshape} differ in shape")
    rank, world = (0, 1)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
    lo, hi = (0, noisy.shape[0]) if presharded else sharding.shard_range(noisy.shape[0], rank, world)
    acc = sharding.MetricAccumulator(dev)
    den_u8 = torch.empty((hi - lo,) + tuple(noisy.shape[1:]), dtype=torch.uint8).pin_memory() if return_denoised else None
    with torch.cuda.device(dev):
        for b0 in range(lo, hi, batch):
            b1 = min(hi, b0 + batch)
            x = noise.u8_to_normalized(noisy[b0:b1].to(dev, non_blocking=True))
            g = noise.u8_to_normalized(gt[b0:b1].to(dev, non_blocking=True))
            den = _denoise(model, x, min(batch, hi - lo))
            psnr, ssim = metrics.batch_metrics(g, den, 2.0)
            acc.update(psnr, ssim)
            if den_u8 is not None:
                den_u8[b0 - lo:b1 - lo].copy_(noise.normalized_to_u8(den), non_blocking=True)
        red = acc.reduce(group)
        torch.cuda.current_stream(dev).synchronize()
    res: Dict[str, object] = {"avg_psnr": red["psnr"], "avg_ssim": red["ssim"], "count": red["count"],
                              "shard": (lo, hi)}
    if den_u8 is not None:
        res["denoised_u8"] = den_u8.numpy()
    return res
