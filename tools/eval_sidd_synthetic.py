"""BASELINE config 4: SIDD-shaped sRGB evaluation on synthetic data.

1280 noisy/clean uint8 block pairs shaped like the SIDD validation .mat files ([40, 32, 256, 256, 3],
evaluate_SIDD/evaluate_SIDD.py:18-41) go through the benchmark.py:32-46 pre/post-processing
(u8 -> [-1,1] -> DiffusionModel(RDUNet_T(32), T=20).improved_sampling -> u8) and the evaluate_SIDD.py:63-64
metrics (PSNR / SSIM, data_range = 2), sharded over the ranks (160 pairs per GPU on 8) with ONE all-reduce of
(sum PSNR, sum SSIM, count).  Launch with torchrun for N > 1.

    python tools/eval_sidd_synthetic.py [--pairs 1280] [--batch 32]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402


def synthetic_blocks(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 256, 256, generator=g)
    x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (4, 4, 4, 4), mode="reflect"), 9, stride=1)
    x = (x - x.amin(dim=(1, 2, 3), keepdim=True)) / (x.amax(dim=(1, 2, 3), keepdim=True) - x.amin(dim=(1, 2, 3), keepdim=True))
    return (x * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1280)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(7)
    model = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20).to(dev).eval()
    if args.precision:
        model.precision = args.precision
    lo, hi = b2.sharding.shard_range(args.pairs, rank, world)
    # the .mat blocks live on the host; each rank touches only its shard ("noisy .mat" = gt + sigma-25 noise, made on
    # the device once, outside the timed region)
    gt_u8 = synthetic_blocks(hi - lo, seed=100 + rank).pin_memory()
    noisy_u8 = torch.empty_like(gt_u8).pin_memory()
    for i0 in range(0, hi - lo, args.batch):
        i1 = min(hi - lo, i0 + args.batch)
        n_u8, _, _ = b2.noise.add_gaussian_noise(gt_u8[i0:i1].to(dev), 25.0, seed=lo + i0, stream_id=3)
        noisy_u8[i0:i1].copy_(n_u8)
    # warm-up (plan + graph capture)
    _ = b2.sidd.denoise_blocks_srgb(model, noisy_u8[:args.batch], batch=args.batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    # benchmark.py:32-46 (u8 -> [-1,1] -> sampler -> u8) + evaluate_SIDD.py:63-64 metrics, batched, one all-reduce
    red = b2.sidd.evaluate_sidd(model, noisy_u8, gt_u8, batch=args.batch, return_denoised=True, presharded=True)
    red = {"count": red["count"], "psnr": red["avg_psnr"], "ssim": red["avg_ssim"]}
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "SIDD-shaped eval (synthetic)", "pairs": red["count"], "n_gpus": world,
                          "precision": model.precision, "mean_psnr_db": red["psnr"], "mean_ssim": red["ssim"],
                          "seconds": float(t[0]), "pairs_per_s": red["count"] / float(t[0]),
                          "ms_per_sample": 1e3 * float(t[0]) * world / red["count"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
