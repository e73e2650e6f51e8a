"""Run the SSIM / PSNR / Welch kernels on a 256-patch batch (the command ncu captures them from)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x, y = (torch.rand(B, 3, 256, 256, device="cuda") for _ in range(2))
for _ in range(3):
    s = b2.metrics.batch_ssim_planes(x.view(-1, 256, 256), y.view(-1, 256, 256), 1.0)
    p = b2.metrics.batch_sse(x, y)
    f, w = b2.metrics.welch(x.flatten(1))
torch.cuda.synchronize()
print("ok", float(s.mean()), float(p.mean()), float(w.mean()))
