"""CPU: host-side partitioning logic, including world_size-2 gloo runs of the metric all-reduce and of the
halo-tiled stitch (the only exchange step of the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from vub_image_denoising_b200 import sharding as sh


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 1280):
        for world in (1, 2, 3, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sh.shard_range(1280, 3, 8) == (480, 640)       # SIDD config: 160 pairs per GPU
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_plan_tiles_4k():
    tiles = sh.plan_tiles(2160, 3840, rows=2, cols=4, halo=200)
    assert len(tiles) == 8
    cover = torch.zeros(2160, 3840, dtype=torch.int32)
    for t in tiles:
        cover[t.y0:t.y1, t.x0:t.x1] += 1
        assert t.y0 % 8 == 0 and t.x0 % 8 == 0 and t.py0 % 8 == 0 and t.px0 % 8 == 0
        assert (t.py1 - t.py0) % 8 == 0 and (t.px1 - t.px0) % 8 == 0
        assert t.py0 == max(0, t.y0 - 200) and t.px1 == min(3840, t.x1 + 200)
    assert int(cover.min()) == 1 and int(cover.max()) == 1
    assert (tiles[0].y1 - tiles[0].y0, tiles[0].x1 - tiles[0].x0) == (1080, 960)
    with pytest.raises(RuntimeError):
        sh.plan_tiles(100, 64, 1, 2)


def _local_net(x):
    """A stand-in 'denoiser' with a 12-px receptive-field radius and zero padding, like the real net's convs."""
    w = torch.ones(3, 1, 5, 5) / 25.0
    y = x
    for _ in range(6):
        y = torch.tanh(F.conv2d(y, w, padding=2, groups=3))
    return y + x


def test_tiled_equals_untiled_single_process():
    g = torch.Generator().manual_seed(0)
    img = torch.rand(1, 3, 96, 160, generator=g)
    full = _local_net(img)
    tiled = sh.denoise_tiled(_local_net, img, rows=2, cols=4, halo=16)
    assert torch.equal(tiled, full)
    # a halo smaller than the receptive field is NOT exact — the halo is what makes tiling valid
    assert not torch.equal(sh.denoise_tiled(_local_net, img, rows=2, cols=4, halo=8), full)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- metrics: each rank owns a contiguous shard of 10 "images"
        psnr_all = torch.arange(10, dtype=torch.float64) + 20
        ssim_all = torch.linspace(0.5, 0.9, 10, dtype=torch.float64)
        bins_all = torch.arange(10) % 5
        lo, hi = sh.shard_range(10, rank, world)
        acc = sh.MetricAccumulator("cpu", n_bins=5)
        acc.update(psnr_all[lo:hi], ssim_all[lo:hi], bins_all[lo:hi])
        red = acc.reduce()
        assert red["count"] == 10
        assert abs(red["psnr"] - float(psnr_all.mean())) < 1e-12
        assert abs(red["ssim"] - float(ssim_all.mean())) < 1e-12
        assert red["count_bins"] == [2] * 5
        assert abs(red["psnr_bins"][1] - float((psnr_all[1] + psnr_all[6]) / 2)) < 1e-12
        # ---- tiled image: 2 ranks, 8 tiles, stitched on rank 0
        g = torch.Generator().manual_seed(0)
        img = torch.rand(1, 3, 96, 160, generator=g)
        out = sh.denoise_tiled(_local_net, img, rows=2, cols=4, halo=16, dst=0)
        if rank == 0:
            assert torch.equal(out, _local_net(img))
        else:
            assert out is None
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: "ok", 1: "ok"}


def _subgroup_worker(rank, world, port, ret):
    """3 processes; the tiled image is denoised by the sub-group {1, 2} and stitched on group rank 1 = global rank 2."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sub = dist.new_group(ranks=[1, 2])
        g = torch.Generator().manual_seed(0)
        img = torch.rand(1, 3, 96, 160, generator=g)
        if rank in (1, 2):
            out = sh.denoise_tiled(_local_net, img, rows=2, cols=4, halo=16, dst=1, group=sub)
            if rank == 2:
                assert torch.equal(out, _local_net(img))
            else:
                assert out is None
            acc = sh.MetricAccumulator("cpu")
            acc.update(torch.tensor([10.0 * rank]), torch.tensor([0.5]))
            red = acc.reduce(group=sub)
            assert red["count"] == 2 and abs(red["psnr"] - 15.0) < 1e-12
            with pytest.raises(ValueError):
                sh.denoise_tiled(_local_net, img, rows=1, cols=2, halo=16, dst=2, group=sub)
        ret[rank] = "ok"
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_tiled_stitch_on_a_subgroup_gloo():
    world = 3
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_subgroup_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: "ok", 1: "ok", 2: "ok"}
