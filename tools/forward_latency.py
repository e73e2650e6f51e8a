"""ms per nn.Module forward call (host-visible latency, synchronised per call) at small batches:
   python tools/forward_latency.py [F=128] [B=1] [reps=20]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
torch.manual_seed(7)
net = b2.RDUNet(base_filters=F).cuda().eval()
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
with torch.no_grad():
    for _ in range(4):
        y = net(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        y = net(x)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
print(f"RDUNet({F}) B={B} 256x256: {dt * 1e3:.3f} ms per forward call (synchronised), {B * 0.065536 / dt:.2f} MPix/s", flush=True)
