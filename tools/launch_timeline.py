"""Where a small-batch forward spends its time between dependent launches (diagnostics build of the library).

    python -m vub_image_denoising_b200._build --timeline          # libb200dn_timeline.so (-DB200DN_TIMELINE)
    B200DN_LIB=vub_image_denoising_b200/libb200dn_timeline.so python tools/launch_timeline.py [B=1] [F=32] [out.txt]

CTA 0 of every tensor-core launch stores %globaltimer at: kernel entry, end of its prologue (barrier init, TMEM
allocation), return of griddepcontrol.wait, first MMA issued, last MMA issued, first accumulator complete (epilogue),
epilogue done, kernel exit.  The script replays one sampler timestep (one forward of 2B images, captured as a CUDA
graph like the product path), dumps the slots of the last replay and prints, per launch, the intervals on one clock."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402
from vub_image_denoising_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/launch_timeline.txt"
L = _lib.lib()
if not hasattr(L, "b200dn_debug_timeline_dump"):
    raise SystemExit("load the diagnostics build: B200DN_LIB=.../libb200dn_timeline.so (python -m vub_image_denoising_b200._build --timeline)")
torch.manual_seed(7)
net = b2.RDUNet_T(base_filters=F).cuda().eval()
net.precision = "fp16"
x = torch.rand(2 * B, 3, 256, 256, device="cuda") * 2 - 1
t = torch.full((2 * B,), 0.5, device="cuda")
with torch.no_grad():
    for _ in range(6):          # call 2 captures the graph, the rest replay it
        y = net(x, t.view(-1, 1, 1, 1))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    for _ in range(10):
        y = net(x, t.view(-1, 1, 1, 1))
e1.record()
torch.cuda.synchronize()
print(f"RDUNet_T({F}) forward of {2 * B} images: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (graph replay + copies)")
raw = out + ".raw"
n = L.b200dn_debug_timeline_dump(C.c_char_p(raw.encode()))
rows = []
for line in open(raw):
    if line.startswith("#"):
        continue
    nums, label = line.split("|")
    v = [int(q) for q in nums.split()]
    if v[1]:
        rows.append((v[0], v[1:10], label.strip()))
# the plan is configured once (69 launches incl. repeated prepare calls are possible): keep the last full set
rows = rows[-56:] if len(rows) >= 56 else rows
t0 = rows[0][1][0]
with open(out, "w") as f:
    hdr = (f"# RDUNet_T({F}) fp16, forward of {2 * B} images 256x256, CUDA-graph replay; CTA 0 of each launch, us on %globaltimer\n"
           "# entry = kernel entry rel. to the first launch; then per launch, relative to its entry: setup (prologue done), dep\n"
           "# (griddepcontrol.wait returned), mma0 (first MMA), mma_end (last MMA issued), acc0 (first accumulator complete),\n"
           "# epi_end, exit; gap = this entry - previous exit (negative: PDL overlap), wait = dep - previous exit\n"
           "#  i   entry   setup     dep    mma0 mma_end    acc0 epi_end    exit |    gap    wait  sm | layer\n")
    f.write(hdr)
    prev_exit = None
    tot = dict(gap=0.0, wait=0.0, fill=0.0, mma=0.0, drain=0.0, n=0)
    for i, v, label in rows:
        rel = [(q - v[0]) / 1e3 if q else float("nan") for q in v[:8]]
        gap = (v[0] - prev_exit) / 1e3 if prev_exit else float("nan")
        wait = (v[2] - prev_exit) / 1e3 if prev_exit and v[2] else float("nan")
        f.write(f"{i:4d} {(v[0] - t0) / 1e3:7.1f} " + " ".join(f"{q:7.2f}" for q in rel[1:]) + f" | {gap:6.2f} {wait:6.2f} {v[8]:3d} | {label}\n")
        prev_exit = v[7] if v[7] else prev_exit
    f.write(f"# first entry -> last exit: {(rows[-1][1][7] - t0) / 1e3:.1f} us for {len(rows)} launches\n")
print(open(out).read())
