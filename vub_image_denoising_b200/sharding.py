"""Multi-GPU partitioning of the hot path — one process per GPU, ``torch.distributed`` plumbing.

The reference is single-device; these are the three shardings BASELINE.json names:

  * independent patches / diffusion samples: contiguous batch split, **no collective** on the data path;
  * metrics: every rank reduces its local (sum PSNR, sum SSIM, count) in fp64 and ONE all-reduce
    (NCCL over NVLink on GPUs, gloo in the CPU tests) combines them;
  * one large image: spatial tiles with a receptive-field halo, each rank denoises its tiles and the
    interiors are stitched on the destination device (the only exchange step of the path).

The receptive-field radius of RDUNet is 193 px, so a 200-px halo (a multiple of 8, which
the three stride-2 levels require) makes the tiled result identical to the untiled one wherever a tile
border is not an image border; at true image borders the network's own zero padding applies.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "allreduce_sums", "MetricAccumulator", "Tile", "plan_tiles", "tiles_of_rank",
           "denoise_tiled", "RF_HALO"]

RF_HALO = 200


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of n independent units for `rank` of `world` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_sums(values: torch.Tensor, group=None) -> torch.Tensor:
    """Sum an fp64 vector over all ranks (identity when not distributed). One collective."""
    if values.dtype != torch.float64:
        raise ValueError("metric sums are reduced in fp64")
    if _dist_on():
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
    return values


class MetricAccumulator:
    """Running (sum PSNR, sum SSIM, count[, per-bin sums]) kept on the device in fp64.

    ``update`` takes per-image device tensors (no host sync); ``reduce`` does the single all-reduce and
    returns python floats (mean PSNR, mean SSIM, count) — the means evaluate_SIDD.py:74-75 prints."""

    def __init__(self, device, n_bins: int = 0):
        self.n_bins = n_bins
        self.acc = torch.zeros(3 + 3 * n_bins, dtype=torch.float64, device=device)

    def update(self, psnr: torch.Tensor, ssim: torch.Tensor, bins: Optional[torch.Tensor] = None) -> None:
        psnr = psnr.to(torch.float64).reshape(-1)
        ssim = ssim.to(torch.float64).reshape(-1)
        self.acc[0] += psnr.sum()
        self.acc[1] += ssim.sum()
        self.acc[2] += psnr.numel()
        if bins is not None and self.n_bins:
            b = bins.to(torch.long).reshape(-1)
            self.acc[3:3 + self.n_bins].index_add_(0, b, psnr)
            self.acc[3 + self.n_bins:3 + 2 * self.n_bins].index_add_(0, b, ssim)
            self.acc[3 + 2 * self.n_bins:].index_add_(0, b, torch.ones_like(psnr))

    def reduce(self, group=None):
        tot = allreduce_sums(self.acc.clone(), group).cpu()
        n = float(tot[2])
        out = {"psnr": float(tot[0]) / n if n else float("nan"),
               "ssim": float(tot[1]) / n if n else float("nan"), "count": int(n)}
        if self.n_bins:
            k = self.n_bins
            cnt = tot[3 + 2 * k:]
            out["psnr_bins"] = (tot[3:3 + k] / cnt).tolist()
            out["ssim_bins"] = (tot[3 + k:3 + 2 * k] / cnt).tolist()
            out["count_bins"] = cnt.to(torch.long).tolist()
        return out


# ------------------------------------------------------------------------------ spatial tiling
@dataclass(frozen=True)
class Tile:
    index: int
    y0: int          # interior (owned) region in the full image
    y1: int
    x0: int
    x1: int
    py0: int         # padded (computed) region = interior + halo, clipped to the image
    py1: int
    px0: int
    px1: int


def _splits(n: int, parts: int, align: int) -> List[int]:
    """`parts`+1 cut points over [0, n], interior cuts multiples of `align`."""
    cuts = [0]
    for i in range(1, parts):
        c = (n * i // parts) // align * align
        cuts.append(max(c, cuts[-1]))
    cuts.append(n)
    return cuts


def plan_tiles(H: int, W: int, rows: int, cols: int, halo: int = RF_HALO, align: int = 8) -> List[Tile]:
    if H % align or W % align:
        raise RuntimeError(f"image size {H}x{W} must be divisible by {align}")
    if halo % align:
        raise ValueError("halo must be a multiple of the alignment")
    ys, xs = _splits(H, rows, align), _splits(W, cols, align)
    tiles = []
    for r in range(rows):
        for c in range(cols):
            y0, y1, x0, x1 = ys[r], ys[r + 1], xs[c], xs[c + 1]
            if y1 <= y0 or x1 <= x0:
                continue
            tiles.append(Tile(len(tiles), y0, y1, x0, x1, max(0, y0 - halo), min(H, y1 + halo),
                              max(0, x0 - halo), min(W, x1 + halo)))
    return tiles


def tiles_of_rank(tiles: Sequence[Tile], rank: int, world: int) -> List[Tile]:
    return [t for t in tiles if t.index % world == rank]


@torch.no_grad()
def denoise_tiled(fn: Callable[[torch.Tensor], torch.Tensor], image: torch.Tensor, rows: int, cols: int,
                  halo: int = RF_HALO, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Denoise one large image ``[B, 3, H, W]`` tile by tile; with torch.distributed initialised the tiles
    are dealt round-robin to the ranks of `group` and the interiors are sent to rank `dst`, which returns the
    stitched image (other ranks return None).  `dst` is a rank WITHIN `group` (== the global rank for the default
    group); the point-to-point calls translate group ranks to the global ranks ``isend`` / ``irecv`` expect.
    Every rank must hold the same input `image` (it is tiny next to the activations: 100 MB for 4K fp32)."""
    B, Cn, H, W = image.shape
    tiles = plan_tiles(H, W, rows, cols, halo)
    world = dist.get_world_size(group) if _dist_on() else 1
    rank = dist.get_rank(group) if _dist_on() else 0
    if rank < 0:
        raise RuntimeError("denoise_tiled: this process is not a member of `group`")
    if not 0 <= dst < world:
        raise ValueError(f"dst={dst} is not a rank of the {world}-rank group")

    def global_rank(r: int) -> int:
        return r if group is None else dist.get_global_rank(group, r)

    mine = tiles_of_rank(tiles, rank, world)
    interiors = {}
    for t in mine:
        crop = image[:, :, t.py0:t.py1, t.px0:t.px1].contiguous()
        out = fn(crop)
        interiors[t.index] = out[:, :, t.y0 - t.py0:t.y1 - t.py0, t.x0 - t.px0:t.x1 - t.px0].contiguous()
    if world == 1:
        full = torch.empty_like(image)
        for t in tiles:
            full[:, :, t.y0:t.y1, t.x0:t.x1] = interiors[t.index]
        return full
    # exchange: point-to-point, interiors land straight in the destination image slices on rank `dst`
    if rank == dst:
        full = torch.empty_like(image)
        pending = []
        for t in tiles:
            owner = t.index % world
            if owner == rank:
                full[:, :, t.y0:t.y1, t.x0:t.x1] = interiors[t.index]
            else:
                buf = torch.empty((B, Cn, t.y1 - t.y0, t.x1 - t.x0), dtype=image.dtype, device=image.device)
                pending.append((t, buf, dist.irecv(buf, src=global_rank(owner), group=group, tag=t.index)))
        for t, buf, work in pending:
            work.wait()
            full[:, :, t.y0:t.y1, t.x0:t.x1] = buf
        return full
    works = [dist.isend(interiors[t.index], dst=global_rank(dst), group=group, tag=t.index) for t in mine]
    for w in works:
        w.wait()
    return None
