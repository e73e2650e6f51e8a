"""Does running a forward as two half-batch launch chains on two streams fill the inter-layer bubbles?
   python tools/two_lane_probe.py F B prec [reps]   (B = total images; compares 1 x B against 2 x B/2)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402
from vub_image_denoising_b200.rdunet import ForwardPlan  # noqa: E402

F, B, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
torch.manual_seed(7)
net = b2.RDUNet(base_filters=F).cuda().eval()
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
out = torch.empty_like(x)
h = B // 2


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    one = ForwardPlan(net, B, 256, 256, prec)
    pa, pb = ForwardPlan(net, h, 256, 256, prec), ForwardPlan(net, h, 256, 256, prec)
    side = torch.cuda.Stream()

    def single():
        one.run(x, out)

    def halves_serial():
        pa.run(x[:h], out[:h])
        pb.run(x[h:], out[h:])

    def lanes():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        pa.run(x[:h], out[:h])
        with torch.cuda.stream(side):
            pb.run(x[h:], out[h:])
        cur.wait_stream(side)

    def graphed(fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay

    ref = None
    for name, fn in (("1 x %d" % B, single), ("2 x %d serial" % h, halves_serial), ("2 x %d two streams" % h, lanes)):
        out.zero_()
        ms = timed(graphed(fn))
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        same = torch.equal(out, ref)
        out.zero_()
        print(f"RDUNet({F}) {prec} {name:22s}: {ms * 1e3:9.1f} us per forward of {B} images (graph replay)  bit-equal to 1 x B: {same}", flush=True)
