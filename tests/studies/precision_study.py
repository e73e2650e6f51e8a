"""GPU: compare every 16-bit precision mode against the stored fp32 oracle outputs (tests/studies/make_study_refs.py)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

REFS = Path(__file__).resolve().parent / "_study_refs"
DEV = "cuda"


def report(tag, got, ref, clean):
    e = (got - ref).abs()
    mse_g = float(((clean - got) ** 2).mean())
    mse_r = float(((clean - ref) ** 2).mean())
    import math
    dpsnr = abs(10 * math.log10(4 / mse_g) - 10 * math.log10(4 / mse_r))
    print(f"  {tag:28s} max {float(e.max()):.3e}  within 1/255: {float((e <= 2 / 255).double().mean()) * 100:8.4f}%  dPSNR {dpsnr:.4f} dB", flush=True)


for f in sorted(REFS.glob("sampler_*.pt")):
    d = torch.load(f)
    torch.manual_seed(d["seed"])
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=32), timesteps=20).to(DEV).eval()
    print(f.name, flush=True)
    for prec in ("bf16", "bf16x2", "fp16", "fp16x2", "bf16x3"):
        dm.precision = prec
        noisy = d["noisy"].to(DEV)
        out = dm.improved_sampling(noisy)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = dm.improved_sampling(noisy)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        report(f"{prec} ({ms:.1f} ms)", out.cpu(), d["ref"], d["clean"])

for f in sorted(REFS.glob("rdunet_*.pt")):
    d = torch.load(f)
    torch.manual_seed(d["seed"])
    net = b2.RDUNet(base_filters=d["F"]).to(DEV).eval()
    print(f.name, flush=True)
    with torch.no_grad():
        for prec in ("bf16", "fp16", "bf16x2", "fp16x2", "bf16x3"):
            net.precision = prec
            out = net(d["noisy"].to(DEV))
            report(prec, out.cpu(), d["ref"], d["clean"])
