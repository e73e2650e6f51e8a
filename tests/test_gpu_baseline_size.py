"""GPU parity at BASELINE.json's full sizes (the round-1 review's holes):

  * config 3 — ``DiffusionModel(RDUNet_T(32), 20).improved_sampling`` on 16 x 256x256, samples checked against the oracle
    (diffusion_denoising/diffusion_RDUnet.py:38-50);
  * config 2 — a sigma = 10 and a sigma = 50 image of the 64-patch sigma-cycling RDUNet(128) batch against the oracle
    (UNet/RDUNet_model.py:157-186);
  * ``direct_sampling`` (diffusion_denoising/diffusion_RDUnet_direct.py:198-201);
  * the fp16 saturation guard, prepared launches == per-call launches, plan invalidation;
  * the NCCL exchange paths (tiled stitch, metric all-reduce) on 2 ranks and a two-device process — skipped on a
    single-GPU box.

Bars (BASELINE.json north_star): >= 99.9 % of pixels within 1/255 in [0,1] space and PSNR within 0.02 dB for the 16-bit
paths; <= 1e-4 max abs error for the fp32 validation build (bf16x3).
"""
import ctypes
import os
import socket
import warnings

import numpy as np
import pytest
import torch

from helpers import check_bar as _check_bar
import vub_image_denoising_b200 as b2
from vub_image_denoising_b200 import _lib
from oracle import rdunet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIGMAS = (10.0, 20.0, 30.0, 40.0, 50.0)


def _image_like_u8(n: int, seed: int) -> np.ndarray:
    """bench.py's synthetic patches: uniform noise low-passed with a 9x9 box, uint8 HWC."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 256, 256, generator=g)
    x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (4, 4, 4, 4), mode="reflect"), 9, stride=1)
    lo, hi = x.amin(dim=(1, 2, 3), keepdim=True), x.amax(dim=(1, 2, 3), keepdim=True)
    return ((x - lo) / (hi - lo) * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()


# ----------------------------------------------------------------------------------------------- config 3
def test_config3_full_size_sampler_vs_oracle(built_lib):
    """B = 16, 256x256, T = 20, RDUNet_T(32), the sampler's default precision: samples 0 and 11 of the batch against
    ``orc.improved_sampling`` (~12 s of CPU each), then the same two samples through the fp32 validation build."""
    torch.manual_seed(7)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20).eval()
    sd = {k: v.clone() for k, v in dm.state_dict().items()}
    dm = dm.to(DEV)
    clean_u8 = torch.from_numpy(_image_like_u8(16, seed=77)).to(DEV)
    _, noisy, clean = b2.noise.add_gaussian_noise(clean_u8, 25.0, seed=5, return_u8=False)
    assert dm.precision == b2.diffusion.SAMPLER_PREC
    out = dm.improved_sampling(noisy)
    assert out.shape == noisy.shape and torch.isfinite(out).all() and not dm.last_saturated
    picks = (0, 11)
    refs = {}
    with torch.no_grad():
        for i in picks:
            refs[i] = orc.improved_sampling(sd, noisy[i:i + 1].cpu(), 20)
    for i in picks:
        _check_bar(out[i:i + 1].cpu(), refs[i], clean[i:i + 1].cpu(), what=f"config 3 sample {i} ({dm.precision})")
    # fp32 validation build on the same two samples (a B = 2 batch: every op of the path is per-sample)
    dm.precision = "bf16x3"
    sub = torch.cat([noisy[i:i + 1] for i in picks]).contiguous()
    val = dm.improved_sampling(sub).cpu()
    for k, i in enumerate(picks):
        mx = float((val[k:k + 1] - refs[i]).abs().max())
        assert mx <= 1e-4, f"config 3 sample {i} bf16x3: max err {mx:.3e}"
    # batch consistency of the default path: the B = 16 rows equal the B = 2 run of the same samples bit for bit
    dm.precision = b2.diffusion.SAMPLER_PREC
    two = dm.improved_sampling(sub)
    for k, i in enumerate(picks):
        assert torch.equal(two[k], out[i])


def test_direct_sampling_vs_oracle(built_lib):
    """f3: the one-shot variant is unet(noisy, t = 1) (diffusion_RDUnet_direct.py:198-201)."""
    torch.manual_seed(19)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20).eval()
    sd = {k: v.clone() for k, v in dm.state_dict().items()}
    dm = dm.to(DEV)
    g = torch.Generator().manual_seed(5)
    clean = torch.rand(2, 3, 64, 96, generator=g) * 2 - 1
    noisy = (clean + torch.randn(2, 3, 64, 96, generator=g) * (25 / 127.5)).clamp(-1, 1)
    with torch.no_grad():
        ref = orc.direct_sampling(sd, noisy)
    keep = noisy.to(DEV)
    for precision, bound in (("bf16", None), ("fp16", None), ("bf16x3", 1e-4)):
        dm.unet.precision = precision
        got = dm.direct_sampling(keep).cpu()
        assert got.shape == ref.shape and got.dtype == torch.float32
        _, mx = _check_bar(got, ref, clean, what=f"direct_sampling {precision}")
        if bound:
            assert mx <= bound
    assert torch.equal(keep.cpu(), noisy), "direct_sampling must not write its input"


# ----------------------------------------------------------------------------------------------- config 2
def test_config2_sigma10_and_sigma50_images_vs_oracle(built_lib):
    """RDUNet(128) bf16 on the 64-patch batch with sigma cycling {10..50}: image 0 (sigma 10) and image 4 (sigma 50) of
    the batch-64 forward against the fp32 oracle, and through the fp32 validation build."""
    torch.manual_seed(7)
    net = b2.RDUNet(base_filters=128).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(DEV)
    clean_u8 = torch.from_numpy(_image_like_u8(64, seed=1234)).to(DEV)
    sigma = torch.tensor([SIGMAS[i % 5] for i in range(64)], device=DEV)
    _, noisy, clean = b2.noise.add_gaussian_noise(clean_u8, sigma, seed=1000, return_u8=False)
    with torch.no_grad():
        out = net(noisy)
        for i in (0, 4):
            assert float(sigma[i]) in (10.0, 50.0)
            ref = orc.rdunet_forward(sd, noisy[i:i + 1].cpu())
            _check_bar(out[i:i + 1].cpu(), ref, clean[i:i + 1].cpu(), what=f"config 2 image {i} sigma {float(sigma[i])}")
            net.precision = "bf16x3"
            val = net(noisy[i:i + 1].contiguous()).cpu()
            net.precision = "bf16"
            mx = float((val - ref).abs().max())
            assert mx <= 1e-4, f"config 2 image {i} bf16x3: max err {mx:.3e}"


# ----------------------------------------------------------------------------------------------- guards
def test_fp16_saturation_guard(built_lib):
    """Weights scaled so that fp16 activations exceed 65504: the launches raise the device flag, improved_sampling warns
    and re-runs the call in the bf16-range fallback mode; sane weights never trip it."""
    torch.manual_seed(5)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=3).to(DEV).eval()
    x = torch.rand(2, 3, 32, 32, device=DEV) * 2 - 1
    dm.precision = "fp16"
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ok = dm.improved_sampling(x)
    assert not dm.last_saturated and torch.isfinite(ok).all()
    # blow the first dense block up: activations ~1e3 x larger per layer
    with torch.no_grad():
        dm.unet.input_block.conv_2.weight.mul_(1e3)
        dm.unet.block_0_0.conv_0.weight.mul_(1e3)
    plan = dm.unet.plan(2, 32, 32, "fp16")
    assert plan.sat_flag is not None and int(plan.sat_flag.item()) == 0
    with torch.no_grad():
        dm.unet.precision = "fp16"
        y16 = dm.unet(x, torch.tensor([0.5], device=DEV).view(1, 1, 1, 1))
    assert int(dm.unet.plan(2, 32, 32, "fp16").sat_flag.item()) == 1, "saturated fp16 stores must raise the flag"
    with pytest.warns(RuntimeWarning, match="saturated"):
        got = dm.improved_sampling(x)
    assert dm.last_saturated
    dm.precision = dm.saturation_fallback
    want = dm.improved_sampling(x)
    assert torch.equal(got, want), "the retry must be the fallback-precision result"
    # bf16 storage has fp32's range: no flag, no retry
    assert dm.unet.plan(2, 32, 32, "bf16").sat_flag is None
    # the guard can be switched off
    dm.precision, dm.check_saturation = "fp16", False
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        dm.improved_sampling(x)
    del y16


def test_prepared_launch_equals_per_call_launch(built_lib):
    """b200dn_igemm_prepare + b200dn_igemm_launch enqueue exactly what b200dn_igemm does (same plan, same maps)."""
    L = _lib.lib()
    torch.manual_seed(0)
    for (B, H, W, cin, cout, ctot, prec) in [(2, 32, 24, 80, 32, 80, _lib.PREC_FP16), (1, 16, 16, 64, 64, 64, _lib.PREC_BF16),
                                             (3, 24, 40, 160, 128, 160, _lib.PREC_BF16)]:
        dt = torch.float16 if prec == _lib.PREC_FP16 else torch.bfloat16
        x = (torch.randn(B, H, W, ctot, device=DEV) * 0.5).to(dt)
        w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
        wp = torch.ops.b200dn.pack_weight(w, prec, False)
        bias = torch.randn(cout, device=DEV) * 0.1
        slope = torch.full((cout,), 0.25, device=DEV)
        outs = [torch.zeros(B, H, W, cout, device=DEV, dtype=dt) for _ in range(3)]
        a = _lib.IgemmArgs()
        a.mode, a.prec, a.B, a.H, a.W, a.cin, a.cout = _lib.MODE_CONV3X3, prec, B, H, W, cin, cout
        a.in_[0], a.in_ctot = x.data_ptr(), ctot
        a.wpacked, a.bias, a.slope = wp.data_ptr(), bias.data_ptr(), slope.data_ptr()
        a.out_kind, a.out_ctot, a.out_coff = _lib.OUT_NHWC16, cout, 0
        st = torch.cuda.current_stream().cuda_stream
        a.out[0] = outs[0].data_ptr()
        _lib.check(L.b200dn_igemm(ctypes.byref(a), st))
        handles = (ctypes.c_void_p * 2)()
        for k in (1, 2):
            a.out[0] = outs[k].data_ptr()
            h = ctypes.c_void_p()
            _lib.check(L.b200dn_igemm_prepare(ctypes.byref(a), ctypes.byref(h)))
            handles[k - 1] = h
        _lib.check(L.b200dn_igemm_launch(handles[0], st))
        _lib.check(L.b200dn_igemm_launch_list(handles, 2, st))      # launches both (the first one again)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]) and float(outs[0].abs().max()) > 0
        assert L.b200dn_igemm_rebind_nchw(handles[0], x.data_ptr(), None, 0) == -1      # not an OUT_NCHW32 launch
        for h in handles:
            L.b200dn_igemm_release(h)


def test_plan_snapshot_and_invalidate(built_lib):
    """A plan snapshots bias / slopes with the packed weights; writes through ``.data`` (no version bump) are picked up
    after ``invalidate_plans()`` — never half (new bias, old weights)."""
    torch.manual_seed(2)
    net = b2.RDUNet(base_filters=16).to(DEV).eval()
    x = torch.rand(1, 3, 16, 16, device=DEV) * 2 - 1
    with torch.no_grad():
        y0 = net(x)
        net.output_block.conv_2.bias.data.add_(1.0)          # bypasses the version counter
        y_stale = net(x)
        assert torch.equal(y_stale, y0), "a cached plan must stay self-consistent"
        net.invalidate_plans()
        y1 = net(x)
    assert float((y1 - y0).abs().min()) > 0.1


def test_forward_from_another_stream(built_lib):
    """The plan's pack kernels run on the stream that built it; a first forward issued from a different stream must be
    ordered after them."""
    torch.manual_seed(4)
    net = b2.RDUNet(base_filters=32).to(DEV).eval()
    x = torch.rand(2, 3, 64, 64, device=DEV) * 2 - 1
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.no_grad():
        with torch.cuda.stream(s):
            plan = net.plan(2, 64, 64)
        out_side = torch.empty_like(x)
        s2 = torch.cuda.Stream()
        with torch.cuda.stream(s2):
            plan.run(x, out_side)
        s2.synchronize()
        ref = net(x)
    assert torch.equal(out_side, ref)


def test_sampler_first_call_inside_user_capture(built_lib):
    """A first improved_sampling call made while the caller captures a CUDA graph must not synchronise or nest a capture."""
    torch.manual_seed(6)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=2).to(DEV).eval()
    x = torch.rand(1, 3, 16, 24, device=DEV) * 2 - 1
    dm.unet.plan(2, 16, 24, dm.precision)          # plan (weight packing) built outside the capture
    torch.cuda.synchronize()
    static_x = x.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = dm.improved_sampling(static_x)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, dm.improved_sampling(x))


# ----------------------------------------------------------------------------------------------- multi-GPU (NCCL)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.manual_seed(1)
        net = b2.RDUNet(base_filters=16).to(dev).eval()
        g = torch.Generator().manual_seed(0)
        img = (torch.rand(1, 3, 512, 768, generator=g) * 2 - 1).to(dev)
        with torch.no_grad():
            tiled = b2.sharding.denoise_tiled(net, img, rows=2, cols=2, dst=0)
            if rank == 0:
                assert torch.equal(tiled, net(img)), "NCCL-stitched tiles differ from the untiled forward"
            else:
                assert tiled is None
        # metric all-reduce: each rank evaluates its shard of 6 image pairs on its own GPU
        a = torch.rand(6, 3, 32, 32, generator=g).to(dev)
        b_ = (a + 0.05 * torch.rand(6, 3, 32, 32, generator=g).to(dev)).clamp(0, 1)
        lo, hi = b2.sharding.shard_range(6, rank, world)
        psnr, ssim = b2.metrics.batch_metrics(a[lo:hi], b_[lo:hi], 1.0)
        acc = b2.sharding.MetricAccumulator(dev)
        acc.update(psnr, ssim)
        red = acc.reduce()
        p_all, s_all = b2.metrics.batch_metrics(a, b_, 1.0)
        assert red["count"] == 6
        assert abs(red["psnr"] - float(p_all.mean())) < 1e-9 and abs(red["ssim"] - float(s_all.mean())) < 1e-9
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL exchange paths)")
def test_two_rank_nccl_tiled_stitch_and_metric_allreduce(built_lib):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: "ok", 1: "ok"}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_process_two_devices(built_lib):
    """The dynamic shared-memory opt-in is per device: a model on cuda:1 after one on cuda:0 must launch."""
    torch.manual_seed(3)
    net = b2.RDUNet(base_filters=32).eval()
    x = torch.rand(1, 3, 64, 64) * 2 - 1
    outs = []
    with torch.no_grad():
        for d in ("cuda:0", "cuda:1"):
            n = b2.RDUNet(base_filters=32).eval()
            n.load_state_dict(net.state_dict())
            outs.append(n.to(d)(x.to(d)).cpu())
    assert torch.equal(outs[0], outs[1])
