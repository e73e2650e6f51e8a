"""Run a few forwards of one configuration (used under ncu for launch lists):
   python tools/forward_once.py rdunet F B [prec] | sampler F B [prec] [T]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

kind, F, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else None
torch.manual_seed(7)
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
with torch.no_grad():
    if kind == "rdunet":
        net = b2.RDUNet(base_filters=F).cuda().eval()
        if prec:
            net.precision = prec
        for _ in range(3):
            y = net(x)
    else:
        T = int(sys.argv[5]) if len(sys.argv) > 5 else 20
        dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=F), timesteps=T).cuda().eval()
        dm.use_cuda_graph = False
        if prec:
            dm.precision = prec
        for _ in range(2):
            y = dm.improved_sampling(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
