"""Device time of one RDUNet_T forward inside a CUDA graph of five back-to-back forwards (events around the replays, no
per-call synchronisation): what a sampler timestep pays.  tools/layer_times.py separates the launches with events; its
sum runs 8 % shorter at F = 32 / 32 images because the gaps let the power-capped part clock higher.
   python tools/forward_graph.py F B prec"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2
F, B, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
torch.manual_seed(7)
net = b2.RDUNet_T(base_filters=F).cuda().eval()
net.precision = prec
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
t = torch.full((1, 1, 1, 1), 0.5, device="cuda")
out = torch.empty_like(x)
plan = net.plan(B, 256, 256)
tt = torch.full((B,), 0.5, device="cuda")
with torch.no_grad():
    for _ in range(3):
        plan.run(x, out, t_ptr=tt.data_ptr(), t_strides=(1, 0, 0))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5):
            plan.run(x, out, t_ptr=tt.data_ptr(), t_strides=(1, 0, 0))
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
print(f"RDUNet_T({F}) {prec} B={B}: {e0.elapsed_time(e1) / 20:.4f} ms per forward (graph of 5 forwards, {1 + len(plan.launches)} launches each)")
