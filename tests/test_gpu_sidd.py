"""GPU tests of the batched SIDD sRGB paths, through the public module and the C ABI.
Reference semantics: evaluate_SIDD/benchmark.py:32-46,79-103 and evaluate_SIDD/evaluate_SIDD.py:43-75."""
import numpy as np
import pytest
import torch

import vub_image_denoising_b200 as b2
from vub_image_denoising_b200 import sidd

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(T=3, F=16, seed=5):
    torch.manual_seed(seed)
    return b2.DiffusionModel(b2.RDUNet_T(base_filters=F), timesteps=T).to(DEV).eval()


def test_batched_blocks_equal_per_block_loop(built_lib):
    """benchmark.py:79-92 visits one block per model call; batches (incl. a ragged, padded last batch) give the same
    uint8 blocks, in the .mat layout's (image, patch) order."""
    dm = _model()
    rng = np.random.default_rng(11)
    blocks = rng.integers(0, 256, size=(2, 3, 32, 40, 3), dtype=np.uint8)      # [I, P, H, W, C]
    out = sidd.denoise_blocks_srgb(dm, blocks, batch=4)                          # 6 blocks: 4 + 2 (padded to 4)
    assert out.shape == blocks.shape and out.dtype == np.uint8
    for i in range(2):
        for j in range(3):
            one = sidd.my_srgb_denoiser(blocks[i, j], dm)
            assert one.shape == blocks[i, j].shape and one.dtype == np.uint8
            assert np.array_equal(one, out[i, j]), f"block ({i},{j}) differs between batch 4 and batch 1"
    # a bare RDUNet (no sampler) goes through forward
    net = b2.RDUNet(base_filters=16).to(DEV).eval()
    o2 = sidd.denoise_blocks_srgb(net, blocks[0], batch=2)
    assert o2.shape == blocks[0].shape


def test_denoiser_stages_vs_oracle(built_lib):
    """Every stage of my_srgb_denoiser against the CPU restatement: ToTensor/Normalize bit-exact, sampler within the
    bf16-path bar, quantiser bit-exact on the same fp32 input."""
    from oracle import noise_oracle as no, rdunet_oracle as orc
    dm = _model()
    sd = {k: v.detach().cpu().clone() for k, v in dm.state_dict().items()}
    rng = np.random.default_rng(12)
    blocks = rng.integers(0, 256, size=(3, 24, 24, 3), dtype=np.uint8)
    x = b2.noise.u8_to_normalized(torch.from_numpy(blocks).to(DEV))
    ref_x = (blocks.astype(np.float32) / np.float32(255.0)).transpose(0, 3, 1, 2)   # ToTensor (benchmark.py:35)
    ref_x = (ref_x - np.float32(0.5)) / np.float32(0.5)                              # Normalize (benchmark.py:36)
    assert np.array_equal(x.cpu().numpy(), ref_x)
    den = dm.improved_sampling(x)
    with torch.no_grad():
        ref_den = orc.improved_sampling(sd, torch.from_numpy(ref_x), 3)
    frac = float(((den.cpu() - ref_den).abs() <= 2 / 255).double().mean())
    assert frac >= 0.999
    out = sidd.denoise_blocks_srgb(dm, blocks, batch=3)
    assert np.array_equal(out, no.norm_to_u8(den.cpu().numpy()))                     # benchmark.py:42-44
    # against the all-CPU chain the uint8 blocks may differ by one code where the sampler differs by < 1/255
    ref_out = no.norm_to_u8(ref_den.numpy())
    assert float((np.abs(out.astype(int) - ref_out.astype(int)) <= 1).mean()) >= 0.999


def test_evaluate_sidd_matches_per_patch_metrics(built_lib):
    """evaluate_SIDD.py:55-75: per-patch PSNR / SSIM with data_range = 2 on channel-last [-1,1] arrays, then the mean."""
    from oracle import metrics_oracle as mo
    dm = _model(seed=6)
    rng = np.random.default_rng(13)
    gt = rng.integers(0, 256, size=(5, 32, 32, 3), dtype=np.uint8)
    noisy = np.clip(gt.astype(np.float32) + rng.normal(0, 25, gt.shape), 0, 255).astype(np.uint8)
    res = sidd.evaluate_sidd(dm, noisy, gt, batch=2, return_denoised=True)
    assert res["count"] == 5 and res["shard"] == (0, 5)
    assert np.array_equal(res["denoised_u8"], sidd.denoise_blocks_srgb(dm, noisy, batch=2))
    # the reference's per-patch loop on the fp32 outputs of the same device path
    x = b2.noise.u8_to_normalized(torch.from_numpy(noisy).to(DEV))
    g = b2.noise.u8_to_normalized(torch.from_numpy(gt).to(DEV)).cpu().numpy()
    den = torch.cat([dm.improved_sampling(x[i:i + 1]) for i in range(5)]).cpu().numpy()
    ps = [mo.peak_signal_noise_ratio(g[i].transpose(1, 2, 0), den[i].transpose(1, 2, 0), data_range=2) for i in range(5)]
    ss = [mo.structural_similarity(g[i].transpose(1, 2, 0), den[i].transpose(1, 2, 0), data_range=2, channel_axis=-1)
          for i in range(5)]
    assert res["avg_psnr"] == pytest.approx(float(np.mean(ps)), abs=1e-4)
    assert res["avg_ssim"] == pytest.approx(float(np.mean(ss)), abs=1e-5)
    with pytest.raises(RuntimeError, match="differ in shape"):
        sidd.evaluate_sidd(dm, noisy, gt[:4])


def test_submission_csv_round_trip_of_device_output(built_lib, tmp_path):
    dm = _model(T=2)
    rng = np.random.default_rng(14)
    blocks = rng.integers(0, 256, size=(1, 2, 16, 16, 3), dtype=np.uint8)
    out = sidd.denoise_blocks_srgb(dm, blocks, batch=2)
    path = tmp_path / "SubmitSrgb.csv"
    assert sidd.write_submission_csv(str(path), out) == 2
    lines = path.read_text().splitlines()
    assert lines[0] == "ID,BLOCK" and len(lines) == 3
    for k, line in enumerate(lines[1:]):
        idx, s = line.split(",", 1)
        assert int(idx) == k
        assert np.array_equal(sidd.base64string_to_array(s, np.uint8, (16, 16, 3)), out[0, k])


def test_benchmark_script_flow_through_the_shim(built_lib, tmp_path):
    """The statement sequence of evaluate_SIDD/benchmark.py (:15-27 import + checkpoint load, :32-46 one-argument
    denoiser, :79-103 block loop + DataFrame.to_csv) on top of the shim, next to the batched path."""
    import pandas as pd
    b2.shim.install()
    try:
        from diffusion_denoising.diffusion_RDUnet import RDUNet_T, DiffusionModel
        torch.manual_seed(3)
        trained = DiffusionModel(RDUNet_T(base_filters=16), timesteps=2)
        ckpt = tmp_path / "diffusion_RDUnet_model_checkpointed_epoch_43.pth"
        torch.save({"epoch": 43, "model_state_dict": trained.state_dict()}, ckpt)       # the reference's checkpoint format
        device = torch.device("cuda")
        model = DiffusionModel(RDUNet_T(base_filters=16), timesteps=2).to(device)
        checkpoint = torch.load(ckpt, map_location=device)
        model.load_state_dict(checkpoint["model_state_dict"])
        model.eval()
        b2.shim.install(model=model)
        rng = np.random.default_rng(2)
        inputs = rng.integers(0, 256, size=(2, 2, 32, 32, 3), dtype=np.uint8)
        strings = []
        for i in range(inputs.shape[0]):
            for j in range(inputs.shape[1]):
                in_block = inputs[i, j, :, :, :]
                out_block = sidd.my_srgb_denoiser(in_block)                                  # one-argument form
                assert in_block.shape == out_block.shape and in_block.dtype == out_block.dtype
                strings.append(sidd.array_to_base64string(out_block))
        df = pd.DataFrame()
        df["ID"] = np.arange(len(strings))
        df["BLOCK"] = strings
        ref_csv = tmp_path / "ref.csv"
        df.to_csv(ref_csv, index=False)
        ours = tmp_path / "SubmitSrgb.csv"
        sidd.write_submission_csv(str(ours), sidd.denoise_blocks_srgb(model, inputs, batch=4))
        assert ours.read_bytes() == ref_csv.read_bytes()
    finally:
        sidd.set_default_model(None)
        b2.shim.uninstall()
    with pytest.raises(RuntimeError, match="no model registered"):
        sidd.my_srgb_denoiser(inputs[0, 0])
