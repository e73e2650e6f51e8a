"""GPU parity of the whole drop-in path against the reference's own outputs (golden vectors produced by
oracle/pin_against_reference.py from the real reference modules) and against the oracle on fresh inputs.

Bars (BASELINE.json north_star): bf16 path >= 99.9 % of pixels within 1/255 in [0,1] space and PSNR within
0.02 dB; fp32 validation build (bf16x3) <= 1e-4 max abs error.
"""
import numpy as np
import pytest
import torch

from helpers import check_bar as _check_bar, sd_digest
import vub_image_denoising_b200 as b2
from oracle import rdunet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("precision,max_abs", [("bf16", None), ("fp16", None), ("bf16x2", None), ("bf16x3", 1e-4)])
def test_rdunet_golden(golden, precision, max_abs, built_lib):
    torch.manual_seed(7)
    net = b2.RDUNet(base_filters=16)
    assert sd_digest(net.state_dict()) == bytes(golden["A_digest"]).hex(), "init differs from the reference ctor"
    net = net.to(DEV).eval()
    net.precision = precision
    with torch.no_grad():
        for k in ("0", "1"):
            x = torch.from_numpy(golden[f"A_x{k}"]).to(DEV)
            ref = torch.from_numpy(golden[f"A_y{k}"]).to(DEV)
            got = net(x)
            assert got.shape == ref.shape and got.dtype == torch.float32
            _check_bar(got, ref, what=f"RDUNet(16) {precision} case {k}")
            if max_abs is not None:
                assert float((got - ref).abs().max()) <= max_abs


@pytest.mark.parametrize("precision,max_abs", [("bf16", None), ("fp16", None), ("bf16x3", 1e-4)])
def test_rdunet_grayscale_golden(golden, precision, max_abs, built_lib):
    """RDUNet(channels=1): the reference ctor takes `channels` for both ends (UNet/RDUNet_model.py:117-155); golden
    vector from the real reference, same seeded init."""
    torch.manual_seed(17)
    net = b2.RDUNet(channels=1, base_filters=16)
    assert sd_digest(net.state_dict()) == bytes(golden["E_digest"]).hex(), "init differs from the reference ctor"
    net = net.to(DEV).eval()
    net.precision = precision
    x = torch.from_numpy(golden["E_x"]).to(DEV)
    ref = torch.from_numpy(golden["E_y"]).to(DEV)
    with torch.no_grad():
        got = net(x)
        assert got.shape == ref.shape == (2, 1, 16, 24)
        _check_bar(got, ref, what=f"grayscale RDUNet(16) {precision}")
        if max_abs is not None:
            assert float((got - ref).abs().max()) <= max_abs
        assert torch.equal(net(x), got)                      # graph replay path
        with pytest.raises(RuntimeError, match="channels"):
            net(torch.zeros(1, 3, 16, 16, device=DEV))
    # grayscale metric path of evaluate_model.calculate_ssim(use_rgb=False)
    from oracle import metrics_oracle as mo
    s = b2.metrics.calculate_ssim(got[0, 0], ref[0, 0], 1.0, use_rgb=False)
    assert s == pytest.approx(mo.structural_similarity(got[0, 0].cpu().numpy(), ref[0, 0].cpu().numpy(), data_range=1.0), abs=1e-5)


@pytest.mark.parametrize("precision,max_abs", [("bf16", None), ("fp16", None), ("bf16x3", 1e-4)])
def test_widths_off_the_channel_tiling_golden(golden, precision, max_abs, built_lib):
    """base_filters = 24 and 10 (the reference ctor takes any width, UNet/RDUNet_model.py:117-155): the plan runs the
    zero-padded 32- / 16-filter embedding; golden vectors from the real reference, same seeded init."""
    with torch.no_grad():
        torch.manual_seed(19)
        net = b2.RDUNet(base_filters=24)
        assert sd_digest(net.state_dict()) == bytes(golden["F_digest24"]).hex()
        net = net.to(DEV).eval()
        net.precision = precision
        x, ref = torch.from_numpy(golden["F_x24"]).to(DEV), torch.from_numpy(golden["F_y24"]).to(DEV)
        got = net(x)
        _check_bar(got, ref, what=f"RDUNet(24) {precision}")
        if max_abs is not None:
            assert float((got - ref).abs().max()) <= max_abs
        assert torch.equal(net(x), got)                      # graph replay path
        assert len(net.state_dict()) == 207 and net.block_0_0.conv_0.weight.shape == (12, 24, 3, 3)
        torch.manual_seed(23)
        net = b2.RDUNet_T(base_filters=10)
        assert sd_digest(net.state_dict()) == bytes(golden["F_digest10"]).hex()
        net = net.to(DEV).eval()
        net.precision = precision
        x, t = torch.from_numpy(golden["F_x10"]).to(DEV), torch.from_numpy(golden["F_t10"]).to(DEV)
        ref = torch.from_numpy(golden["F_y10"]).to(DEV)
        got = net(x, t)
        _check_bar(got, ref, what=f"RDUNet_T(10) {precision}")
        if max_abs is not None:
            assert float((got - ref).abs().max()) <= max_abs
        # a parameter update is picked up (the padded copy is rebuilt with the plan)
        net.output_block.conv_2.weight.zero_()
        net.output_block.conv_2.bias.zero_()
        assert torch.equal(net(x, t), x)                     # PReLU(0) + global residual


@pytest.mark.parametrize("precision,max_abs", [("bf16", None), ("bf16x2", None), ("bf16x3", 1e-4)])
def test_rdunet_t_golden(golden, precision, max_abs, built_lib):
    torch.manual_seed(11)
    net = b2.RDUNet_T(base_filters=16)
    assert sd_digest(net.state_dict()) == bytes(golden["B_digest"]).hex()
    net = net.to(DEV).eval()
    net.precision = precision
    x = torch.from_numpy(golden["B_x"]).to(DEV)
    with torch.no_grad():
        for k in ("0", "1"):
            t = torch.from_numpy(golden[f"B_t{k}"]).to(DEV)
            ref = torch.from_numpy(golden[f"B_y{k}"]).to(DEV)
            got = net(x, t)
            _check_bar(got, ref, what=f"RDUNet_T(16) {precision} t{k}")
            if max_abs is not None:
                assert float((got - ref).abs().max()) <= max_abs


@pytest.mark.parametrize("use_graph", [True, False])
@pytest.mark.parametrize("precision,max_abs", [("fp16", None), ("fp16x2", None), ("bf16x2", None), ("bf16x3", 1e-4)])
def test_sampler_golden(golden, precision, max_abs, use_graph, built_lib):
    torch.manual_seed(13)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=4)
    assert sd_digest(dm.state_dict()) == bytes(golden["C_digest"]).hex()
    dm = dm.to(DEV).eval()
    dm.precision = precision
    dm.use_cuda_graph = use_graph
    noisy = torch.from_numpy(golden["C_noisy"]).to(DEV)
    clean = torch.from_numpy(golden["C_clean"]).to(DEV)
    ref = torch.from_numpy(golden["C_out"]).to(DEV)
    keep = noisy.clone()
    got = dm.improved_sampling(noisy)
    assert torch.equal(noisy, keep), "improved_sampling must not write its input"
    assert got.data_ptr() != noisy.data_ptr()
    _check_bar(got, ref, what=f"sampler T=4 {precision}")
    if max_abs is not None:
        assert float((got - ref).abs().max()) <= max_abs
    # second call reuses the captured graph / cached plan and must give the same bits
    again = dm.improved_sampling(noisy)
    assert torch.equal(got, again)
    # forward_diffusion is exact fp32 arithmetic
    fd = dm.forward_diffusion(clean, noisy, 3)
    assert torch.equal(fd, torch.from_numpy(golden["C_fd3"]).to(DEV))
    # forward = forward_diffusion + improved_sampling
    full = dm(clean, noisy, 3)
    assert torch.equal(full, dm.improved_sampling(fd))


def test_eval_width_t32_golden(golden, built_lib):
    """The reference's evaluation width for the sampler network: RDUNet_T(base_filters=32)."""
    torch.manual_seed(7)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20)
    assert sd_digest(dm.state_dict()) == bytes(golden["D_digest_T32"]).hex()
    dm = dm.to(DEV).eval()
    x = torch.from_numpy(golden["D_x"]).to(DEV)
    ref = torch.from_numpy(golden["D_y"]).to(DEV)
    t = torch.tensor([0.5], device=DEV).view(1, 1, 1, 1)
    with torch.no_grad():
        for precision, bound in (("bf16", None), ("bf16x2", None), ("bf16x3", 1e-4)):
            dm.unet.precision = precision
            got = dm.unet(x, t)
            frac, mx = _check_bar(got, ref, what=f"RDUNet_T(32) {precision}")
            if bound:
                assert mx <= bound, f"{precision}: max err {mx:.3e}"


def test_sampler_full_schedule_vs_oracle(built_lib):
    """Full 20-step schedule, F=32, against the fp32 CPU oracle on a fresh seeded input (64x64 keeps the oracle
    to a few seconds).  The sampler's default 16-bit mode and the fp16 modes must hold the 99.9%-within-1/255
    bar, bf16x3 the 1e-4 bar.  (bf16 WEIGHT rounding alone breaks the bar on this seed — 99.79% — on the GPU
    and in a CPU emulation alike, which is why the sampler does not default to a bf16 mode)"""
    torch.manual_seed(3)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20).eval()
    g = torch.Generator().manual_seed(99)
    clean = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    noisy = (clean + torch.randn(2, 3, 64, 64, generator=g) * (25 / 127.5)).clamp(-1, 1)
    with torch.no_grad():
        ref = orc.improved_sampling(dm.state_dict(), noisy, 20)
    dm = dm.to(DEV)
    assert dm.precision == b2.diffusion.SAMPLER_PREC
    for precision, bound in ((dm.precision, None), ("fp16", None), ("fp16x2", None), ("bf16x3", 1e-4)):
        dm.precision = precision
        got = dm.improved_sampling(noisy.to(DEV)).cpu()
        frac, mx = _check_bar(got, ref, clean, what=f"20-step sampler {precision}")
        if bound:
            assert mx <= bound, f"{precision}: max err {mx:.3e}"


def test_rdunet128_full_size_patch_vs_oracle(built_lib):
    """BASELINE config 1: RDUNet(base_filters=128), one 256x256 RGB patch, sigma = 25, fp32 CPU oracle vs the
    bf16 GPU path (and the fp32-validation build)."""
    torch.manual_seed(7)
    net = b2.RDUNet(base_filters=128).eval()
    rng = np.random.default_rng(0)
    clean_u8 = torch.from_numpy(rng.integers(0, 256, size=(1, 256, 256, 3), dtype=np.uint8))
    _, noisy, clean = b2.noise.add_gaussian_noise(clean_u8.to(DEV), 25.0, seed=1)
    with torch.no_grad():
        ref = orc.rdunet_forward(net.state_dict(), noisy.cpu())
        net = net.to(DEV)
        for precision, bound in (("bf16", None), ("bf16x3", 1e-4)):
            net.precision = precision
            got = net(noisy).cpu()
            frac, mx = _check_bar(got, ref, clean.cpu(), what=f"RDUNet(128) 256x256 {precision}")
            if bound:
                assert mx <= bound, f"{precision}: max err {mx:.3e}"


def test_config2_full_batch_is_batch_consistent(built_lib):
    """BASELINE config 2 at full size (RDUNet(128), 64 x 256x256, sigma cycling 10..50): a size-independent property.
    Every op of the network is per-sample, so image i of the batch-64 forward (CTA-pair kernels, MT = 2 tiles, full
    grids) must equal the batch-1 forward of that image (under-filled grids, split N tiles, different kernels) bit for
    bit, and the on-device PSNR/SSIM of the batch must equal the per-image values."""
    from oracle import metrics_oracle as mo
    torch.manual_seed(7)
    net = b2.RDUNet(base_filters=128).to(DEV).eval()
    rng = np.random.default_rng(21)
    clean_u8 = rng.integers(0, 256, size=(64, 256, 256, 3), dtype=np.uint8)
    sigma = torch.tensor([(10.0, 20.0, 30.0, 40.0, 50.0)[i % 5] for i in range(64)], device=DEV)
    _, noisy, clean = b2.noise.add_gaussian_noise(torch.from_numpy(clean_u8).to(DEV), sigma, seed=3, return_u8=False)
    with torch.no_grad():
        out = net(noisy)
        assert torch.isfinite(out).all()
        for i in (0, 17, 63):
            one = net(noisy[i:i + 1].contiguous())
            assert torch.equal(one[0], out[i]), f"image {i}: batch-64 and batch-1 forwards differ"
    psnr, ssim = b2.metrics.batch_metrics(clean, out, 1.0)
    for i in (0, 63):
        c, o = clean[i].cpu().numpy(), out[i].cpu().numpy()
        assert float(psnr[i]) == pytest.approx(mo.calculate_psnr(c, o, 1.0), abs=1e-4)
        assert float(ssim[i]) == pytest.approx(mo.structural_similarity(c, o, data_range=1.0, channel_axis=0), abs=1e-5)


def test_forward_graph_replay_and_user_capture(built_lib):
    """The nn.Module forward replays its own CUDA graph from the second call on; a caller that captures the forward
    into a graph of its own gets the eager launch list inside that capture.  All three give the same bits, for changing
    inputs, and for RDUNet_T with scalar and per-sample timesteps."""
    torch.manual_seed(3)
    net = b2.RDUNet(base_filters=16).to(DEV).eval()
    xs = [torch.rand(2, 3, 32, 48, device=DEV) * 2 - 1 for _ in range(3)]
    with torch.no_grad():
        first = net(xs[0])                      # call 1: eager
        again = net(xs[0])                      # call 2: captures + replays
        assert torch.equal(first, again)
        outs = [net(x) for x in xs]             # replays with new inputs
        plan = net.plan(2, 32, 48)
        assert plan._graph is not None
        ref = [torch.empty_like(x) for x in xs]
        for x, r in zip(xs, ref):
            plan.run(x, r)                      # eager launch list
        torch.cuda.synchronize()
        for o, r in zip(outs, ref):
            assert torch.equal(o, r)
        assert outs[0].data_ptr() != outs[1].data_ptr()      # fresh tensors, not views of the static buffer
        # user-side capture
        static_x = xs[1].clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(static_x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = net(static_x)
        static_x.copy_(xs[2])
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, ref[2])
        # RDUNet_T: scalar and per-sample timesteps through the graph path
        nt = b2.RDUNet_T(base_filters=16).to(DEV).eval()
        t1 = torch.tensor([0.35], device=DEV).view(1, 1, 1, 1)
        tb = torch.tensor([0.2, 0.9], device=DEV).view(2, 1, 1, 1)
        a1, a2 = nt(xs[0], t1), nt(xs[0], t1)
        b1_ = nt(xs[0], tb)
        assert torch.equal(a1, a2)
        one0 = nt(xs[0][:1].contiguous(), tb[:1])
        one1 = nt(xs[0][1:].contiguous(), tb[1:])
        assert torch.equal(b1_[0], one0[0]) and torch.equal(b1_[1], one1[0])


def test_tiled_image_equals_untiled(built_lib):
    """BASELINE config 5 in small: halo-200 tiling (receptive-field radius 193) reproduces the untiled forward
    bit for bit; the only zero padding is at true image borders."""
    torch.manual_seed(1)
    net = b2.RDUNet(base_filters=16).to(DEV).eval()
    img = torch.rand(1, 3, 512, 768, device=DEV) * 2 - 1
    with torch.no_grad():
        full = net(img)
        tiled = b2.sharding.denoise_tiled(net, img, rows=2, cols=2)
        assert torch.equal(full, tiled)
        small_halo = b2.sharding.denoise_tiled(net, img, rows=2, cols=2, halo=64)
        assert not torch.equal(full, small_halo)


def test_sidd_shaped_pipeline_vs_oracle(built_lib):
    """BASELINE config 4 in small: u8 blocks -> normalise -> sampler -> quantise + PSNR/SSIM (R = 2), against the
    oracle's arithmetic for every stage (benchmark.py:32-46, evaluate_SIDD.py:63-64)."""
    from oracle import metrics_oracle as mo, noise_oracle as no
    torch.manual_seed(5)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=3).eval()
    sd = {k: v.clone() for k, v in dm.state_dict().items()}
    dm = dm.to(DEV)
    rng = np.random.default_rng(3)
    gt_u8 = rng.integers(0, 256, size=(2, 32, 32, 3), dtype=np.uint8)
    noisy_u8, _, clean = b2.noise.add_gaussian_noise(torch.from_numpy(gt_u8).to(DEV), 25.0, seed=9)
    noisy = b2.noise.u8_to_normalized(noisy_u8)
    den = dm.improved_sampling(noisy)
    out_u8 = b2.noise.normalized_to_u8(den)
    psnr, ssim = b2.metrics.batch_metrics(clean, den, 2.0)
    # oracle chain on the CPU
    r_noisy_u8, r_noisy, r_clean = no.degrade(gt_u8, 25.0, seed=9)
    assert np.array_equal(noisy_u8.cpu().numpy(), r_noisy_u8) and np.array_equal(noisy.cpu().numpy(), r_noisy)
    with torch.no_grad():
        r_den = orc.improved_sampling(sd, torch.from_numpy(r_noisy), 3)
    _check_bar(den.cpu(), r_den, what="SIDD-shaped sampler")
    # quantiser is exact on the same input
    assert np.array_equal(out_u8.cpu().numpy(), no.norm_to_u8(den.cpu().numpy()))
    for i in range(2):
        g, o = r_clean[i].transpose(1, 2, 0), den[i].cpu().numpy().transpose(1, 2, 0)
        assert float(psnr[i]) == pytest.approx(mo.peak_signal_noise_ratio(g, o, data_range=2), abs=1e-4)
        assert float(ssim[i]) == pytest.approx(mo.structural_similarity(g, o, data_range=2, channel_axis=-1), abs=1e-5)


def test_module_contract(built_lib):
    net = b2.RDUNet(base_filters=16).to(DEV).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="divisible by 8"):
            net(torch.zeros(1, 3, 20, 32, device=DEV))
        with pytest.raises(RuntimeError, match="channels"):
            net(torch.zeros(1, 4, 16, 16, device=DEV))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            net(torch.zeros(1, 3, 16, 16))
        x = torch.rand(2, 3, 16, 16, device=DEV)
        y0 = net(x)
        # load_state_dict must invalidate packed weights
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        sd["output_block.conv_2.bias"] += 1.0
        net.load_state_dict(sd)
        y1 = net(x)
        assert float((y1 - y0).abs().min()) > 0.1
        for k in sd:
            if k.endswith("conv_2.weight") and k.startswith("output_block"):
                sd[k] = sd[k] * 0
        net.load_state_dict(sd)
        y2 = net(x)
        # zero output conv -> out = prelu(bias) + x
        b = sd["output_block.conv_2.bias"].to(DEV)
        a = sd["output_block.actv_2.weight"].to(DEV)
        expect = torch.where(b > 0, b, a * b).view(1, 3, 1, 1) + x
        assert float((y2 - expect).abs().max()) < 1e-6
    with pytest.raises(RuntimeError, match="inference-only"):
        net.train()(torch.rand(1, 3, 16, 16, device=DEV))


def test_linearity_of_sampler_step_and_idempotent_shapes(built_lib):
    """Size-independent property at BASELINE size: zero networks make the sampler an exact identity chain."""
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=20).to(DEV).eval()
    with torch.no_grad():
        for p in dm.parameters():
            p.zero_()
    y = torch.rand(16, 3, 256, 256, device=DEV) * 2 - 1
    out = dm.improved_sampling(y)
    # U(x,t) = 0 + x  =>  x_{t-1} = x - ((1-a_t) x + a_t y) + ((1-a_p) x + a_p y); with x_T = y every step returns y
    # up to fp32 rounding of the reference's operation order — which the oracle reproduces exactly.
    x = y.cpu()
    for t in range(20, 0, -1):
        x = orc.sampler_step(x, x, x, y.cpu(), t, 20)
    assert torch.equal(out.cpu(), x)


@pytest.mark.parametrize("base_filters,shape", [(16, (1, 3, 8, 8)), (16, (3, 3, 8, 24)), (48, (1, 3, 40, 56)), (32, (5, 3, 24, 16)),
                                                  (16, (2, 3, 136, 72))])
def test_ragged_shapes_and_widths_vs_oracle(base_filters, shape, built_lib):
    """Edge cases: the smallest legal image (8x8 -> 1x1 at the deepest level), non-power-of-two widths
    (base_filters = 48 -> 24-channel growth, N padded to 32), odd batches, sizes that are not multiples of the
    8x16 / 16x8 accumulator tiles."""
    torch.manual_seed(base_filters + shape[2])
    net = b2.RDUNet(base_filters=base_filters).eval()
    g = torch.Generator().manual_seed(4)
    x = torch.rand(*shape, generator=g) * 2 - 1
    with torch.no_grad():
        ref = orc.rdunet_forward(net.state_dict(), x)
        net = net.to(DEV)
        net.precision = "bf16x3"
        got = net(x.to(DEV)).cpu()
        assert float((got - ref).abs().max()) <= 1e-4
        net.precision = "fp16"
        got = net(x.to(DEV)).cpu()
        _check_bar(got, ref, what=f"RDUNet({base_filters}) {shape} fp16")


def test_unsupported_configurations_raise(built_lib):
    with torch.no_grad():
        two = b2.RDUNet(channels=2, base_filters=16).to(DEV).eval()
        with pytest.raises(RuntimeError, match="RGB .3-channel. and grayscale"):
            two(torch.zeros(1, 2, 16, 16, device=DEV))
        net = b2.RDUNet(base_filters=16).to(DEV).eval()
        net.precision = "int8"
        with pytest.raises(RuntimeError, match="unknown precision"):
            net(torch.zeros(1, 3, 16, 16, device=DEV))
