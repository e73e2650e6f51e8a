"""Small end-to-end run for compute-sanitizer: every kernel family once at tiny sizes."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
rng = np.random.default_rng(0)
clean_u8 = torch.from_numpy(rng.integers(0, 256, size=(2, 24, 40, 3), dtype=np.uint8)).to(dev)
_, noisy, clean = b2.noise.add_gaussian_noise(clean_u8, [10.0, 50.0], seed=3)
with torch.no_grad():
    for F, prec in ((16, "bf16"), (16, "fp16x2"), (32, "bf16x3")):
        net = b2.RDUNet(base_filters=F).to(dev).eval()
        net.precision = prec
        y = net(noisy)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=2).to(dev).eval()
    dm.use_cuda_graph = False
    z = dm.improved_sampling(noisy)
    psnr, ssim = b2.metrics.batch_metrics(clean, z, 2.0)
    u8 = b2.noise.normalized_to_u8(z)
    back = b2.noise.u8_to_normalized(u8)
torch.cuda.synchronize()
print("ok", float(psnr.mean()), float(ssim.mean()), tuple(u8.shape), float(back.abs().max()))
