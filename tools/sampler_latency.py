"""ms per improved_sampling call at a small batch (config 3 on 8 GPUs runs 2 samples per GPU):
   python tools/sampler_latency.py [B=2] [F=32] [T=20] [reps=5]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
torch.manual_seed(7)
dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=F), timesteps=T).cuda().eval()
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
for _ in range(3):
    dm.improved_sampling(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    dm.improved_sampling(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"improved_sampling B={B} F={F} T={T} {dm.precision}: {ms:.3f} ms per call, {ms / B:.3f} ms per sample, "
      f"{ms / T:.4f} ms per timestep (one forward of {2 * B} images + update)", flush=True)
