"""Is the transposed conv limited by its 2x2 scatter?  Same GEMM (M = B*H*W pixels, N = 4*C, K = C) once as UP2X2
(phase tiles scattered to (2y+ky, 2x+kx) of a [B,2H,2W,1.5C] buffer) and once as a 1x1 conv writing [B,H,W,4C]
contiguously.  python tools/up_vs_1x1.py [C=128] [B=64] [H=128]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vub_image_denoising_b200 as b2  # noqa: E402
from vub_image_denoising_b200 import _lib  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
prec = _lib.PREC_BF16
dev = "cuda"
x = (torch.randn(B, H, H, C * 5 // 2, device=dev) * 0.5).to(torch.bfloat16).view(torch.int16)
wt = torch.randn(C, C, 2, 2, device=dev) * 0.05            # ConvTranspose2d weight [Cin, Cout, 2, 2]
w1 = torch.randn(4 * C, C, 1, 1, device=dev) * 0.05         # 1x1 conv with N = 4C
bias = torch.zeros(C, device=dev)
bias4 = torch.zeros(4 * C, device=dev)
slope = torch.full((C,), 0.25, device=dev)
slope4 = torch.full((4 * C,), 0.25, device=dev)
wp_t = torch.ops.b200dn.pack_weight(wt, prec, True)
wp_1 = torch.ops.b200dn.pack_weight(w1, prec, False)
out_up = torch.empty((B, 2 * H, 2 * H, C // 2 + C), dtype=torch.int16, device=dev)
out_1 = torch.empty((B, H, H, 4 * C), dtype=torch.int16, device=dev)


def run(fn, name, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * H * C * 4 * C
    print(f"{name:34s} {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  output {nbytes / ms / 1e6:7.1f} GB/s", flush=True)


nb = B * 4 * H * H * C * 2
run(lambda: torch.ops.b200dn.conv_igemm(x, None, wp_t, bias, slope, _lib.MODE_UP2X2, prec, C, C, out_up, None, C // 2, None, None),
    f"UP2X2 C={C} B={B} {H}x{H} (scatter)", nb)
run(lambda: torch.ops.b200dn.conv_igemm(x, None, wp_1, bias4, slope4, _lib.MODE_CONV1X1, prec, C, 4 * C, out_1, None, 0, None, None),
    f"1x1 N=4C (contiguous rows)", nb)
