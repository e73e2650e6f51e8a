"""CPU (no GPU needed): fp32 oracle outputs of the 20-step sampler / RDUNet forward for several seeds, stored under
tests/studies/_study_refs/ (git-ignored, travels with the gpurun snapshot) so the GPU box only has to run the CUDA side."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402
from oracle import rdunet_oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent / "_study_refs"
OUT.mkdir(exist_ok=True)
torch.set_grad_enabled(False)


def inputs(seed, B, hw):
    g = torch.Generator().manual_seed(1000 + seed)
    clean = torch.rand(B, 3, hw, hw, generator=g)
    clean = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(clean, (4, 4, 4, 4), mode="reflect"), 9, stride=1)
    clean = clean * 2 - 1
    noisy = (clean + torch.randn(B, 3, hw, hw, generator=g) * (25 / 127.5)).clamp(-1, 1)
    return clean, noisy


for seed, B, hw in [(3, 2, 64), (7, 2, 64), (11, 2, 64), (21, 2, 64), (33, 2, 64), (5, 1, 256)]:
    f = OUT / f"sampler_f32_s{seed}_b{B}_{hw}.pt"
    if f.exists():
        continue
    torch.manual_seed(seed)
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20).eval()
    clean, noisy = inputs(seed, B, hw)
    ref = orc.improved_sampling(dm.state_dict(), noisy, 20)
    torch.save({"seed": seed, "clean": clean, "noisy": noisy, "ref": ref}, f)
    print("wrote", f.name, flush=True)

for seed, F, B, hw in [(7, 128, 1, 256), (9, 128, 1, 256), (7, 32, 2, 256), (7, 64, 1, 256)]:
    f = OUT / f"rdunet_f{F}_s{seed}_b{B}_{hw}.pt"
    if f.exists():
        continue
    torch.manual_seed(seed)
    net = b2.RDUNet(base_filters=F).eval()
    clean, noisy = inputs(seed, B, hw)
    ref = orc.rdunet_forward(net.state_dict(), noisy)
    torch.save({"seed": seed, "F": F, "clean": clean, "noisy": noisy, "ref": ref}, f)
    print("wrote", f.name, flush=True)
