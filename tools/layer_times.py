"""Per-launch device times of the forward inside a sustained loop (power-capped clocks), by event pairs.
   python tools/layer_times.py [F=128] [B=64] [prec=bf16] [reps=8]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
torch.manual_seed(7)
net = b2.RDUNet(base_filters=F).cuda().eval()
net.precision = prec
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
out = torch.empty_like(x)
plan = net.plan(B, 256, 256)
n = len(plan.launches)
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(reps)]
with torch.no_grad():
    for _ in range(4):
        plan.run(x, out)
    for r in range(reps):
        plan.run(x, out, layer_events=evs[r])
torch.cuda.synchronize()
tot = 0.0
groups = {}
print(f"RDUNet({F}) {prec} B={B}: per tensor-core launch, mean of {reps} sustained forwards")
for i, info in enumerate(plan.layer_info):
    ms = sum(evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(reps)) / reps
    tot += ms
    tf = info["flops"] / ms / 1e9
    kind = {0: "conv3x3", 1: "down", 2: "up", 4: "denseblk", 5: "chain4"}[info["mode"]]
    key = f"{kind} N={min(info['cout'], 256) if kind != 'up' else 256}"
    g = groups.setdefault(key, [0.0, 0.0])
    g[0] += ms
    g[1] += info["flops"]
    print(f"  {i:2d} {kind:8s} {info['H']:4d}x{info['W']:<4d} cin {info['cin']:5d} cout {info['cout']:5d}  {ms:8.3f} ms  {tf:8.1f} TFLOP/s")
print(f"total {tot:.2f} ms, {plan.flops / tot / 1e9:.1f} TFLOP/s (all tensor-core launches)")
for k, (ms, fl) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:16s} {ms:8.2f} ms {ms / tot * 100:5.1f}%  {fl / ms / 1e9:8.1f} TFLOP/s")
