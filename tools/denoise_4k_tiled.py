"""BASELINE config 5: one 3840x2160 RGB image through RDUNet(128), split into 2 x 4 spatial tiles with a
200-px receptive-field halo, tiles dealt to the ranks, interiors stitched on the destination device.

    python tools/denoise_4k_tiled.py [--check]     (torchrun for N > 1; --check also runs the untiled image on rank 0)
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--base-filters", type=int, default=128)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(7)
    net = b2.RDUNet(base_filters=args.base_filters).to(dev).eval()
    g = torch.Generator().manual_seed(5)
    img = (torch.rand(1, 3, args.height, args.width, generator=g) * 2 - 1).to(dev)    # every rank holds the input
    with torch.no_grad():
        out = b2.sharding.denoise_tiled(net, img, rows=2, cols=4)                       # warm-up (plans)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = b2.sharding.denoise_tiled(net, img, rows=2, cols=4)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res = {"config": "4K tiled", "n_gpus": world, "height": args.height, "width": args.width,
               "tiles": 8, "halo": b2.sharding.RF_HALO, "seconds": dt, "mpix_per_s": args.height * args.width / 1e6 / dt}
        if args.check and rank == 0:
            full = net(img)
            res["max_abs_diff_vs_untiled"] = float((full - out).abs().max())
            res["bit_identical"] = bool(torch.equal(full, out))
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
