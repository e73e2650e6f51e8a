"""Latency of the nn.Module forward at small batches, and the HOST cost of enqueueing one eager (non-graph) forward:
   python tools/forward_latency.py [F=128] [B=1] [reps=20]
  * ms per forward call (synchronised per call; graph replay from the second call on);
  * host microseconds per launch of an eager forward: wall time for ForwardPlan.run to return (no synchronisation
    inside the loop), divided by the launches it enqueues (1 ingest + 68 prepared tcgen05 launches)."""
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
    plan = net.plan(B, 256, 256)
    out = torch.empty_like(x)
    n_launch = 1 + len(plan.launches)
    host = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            plan.run(x, out)
        host.append((time.perf_counter() - t0) / 4)
        torch.cuda.synchronize()
    h = min(host)
    print(f"eager ForwardPlan.run host time: {h * 1e6:.1f} us per forward = {h * 1e6 / n_launch:.2f} us per launch "
          f"({n_launch} launches, prepared tensor maps, one C call for the 68 igemm launches)", flush=True)
