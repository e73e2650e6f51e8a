"""Run ONE tensor-core layer a few times (the command ncu captures per-kernel counters from):
   python tools/ncu_layer.py conv H cin cout ctot [B=32] [prec=fp16] [res=0] [out_ctot=ctot] [out_coff=cin]
   python tools/ncu_layer.py up   H cin cout      [B=32] [prec=fp16]      (transposed 2x2: input HxH, output 2Hx2H)
   python tools/ncu_layer.py down H cin cout      [B=32] [prec=fp16]
Prints the event-timed mean so the plain run's number sits beside the profile."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vub_image_denoising_b200 import _lib  # noqa: E402

kind = sys.argv[1]
H, cin, cout = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rest = sys.argv[5:]
if kind == "conv":
    ctot = int(rest[0])
    rest = rest[1:]
else:
    ctot = cin
B = int(rest[0]) if len(rest) > 0 else 32
prec_name = rest[1] if len(rest) > 1 else "fp16"
res = int(rest[2]) if len(rest) > 2 else 0
out_ctot = int(rest[3]) if len(rest) > 3 else (ctot if kind == "conv" else (cout if kind == "down" else 3 * cout // 2))
out_coff = int(rest[4]) if len(rest) > 4 else (min(cin, out_ctot - cout) if kind == "conv" else 0)
prec = _lib.PREC_NAMES[prec_name]
dt = torch.float16 if prec in _lib.FP16_PRECS else torch.bfloat16
dev = "cuda"
L = _lib.lib()
torch.manual_seed(0)
x = (torch.randn(B, H, H, ctot, device=dev) * 0.5).to(dt)
mode = {"conv": _lib.MODE_CONV3X3, "up": _lib.MODE_UP2X2, "down": _lib.MODE_DOWN2X2}[kind]
if kind == "up":
    w = torch.randn(cin, cout, 2, 2, device=dev) * 0.05
    groups, Ho = 4, 2 * H
elif kind == "down":
    w = torch.randn(cout, cin, 2, 2, device=dev) * 0.05
    groups, Ho = 4, H // 2
else:
    w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    groups, Ho = 9, H
nbytes = L.b200dn_packed_weight_bytes(cout, cin, groups, prec)
wp = torch.empty(nbytes // 2, dtype=torch.int16, device=dev)
st = torch.cuda.current_stream().cuda_stream
if kind == "up":
    _lib.check(L.b200dn_pack_convt_weight(w.data_ptr(), cin, cout, prec, wp.data_ptr(), st))
else:
    _lib.check(L.b200dn_pack_conv_weight(w.data_ptr(), cout, cin, w.shape[2], w.shape[3], prec, wp.data_ptr(), st))
bias = torch.zeros(cout, device=dev)
slope = torch.full((cout,), 0.25, device=dev)
out = x if (kind == "conv" and out_ctot == ctot and out_coff >= cin) else torch.empty(B, Ho, Ho, out_ctot, device=dev, dtype=dt)
a = _lib.IgemmArgs()
a.mode, a.prec, a.B, a.H, a.W, a.cin, a.cout = mode, prec, B, H, H, cin, cout
a.in_[0] = x.data_ptr()
a.in_ctot = ctot
a.wpacked, a.bias, a.slope = wp.data_ptr(), bias.data_ptr(), slope.data_ptr()
a.out_kind = _lib.OUT_NHWC16
a.out[0] = out.data_ptr()
a.out_ctot, a.out_coff = out_ctot, out_coff
import os
a.m_tiles = int(os.environ.get("M_TILES", "0"))
a.block_n = int(os.environ.get("BLOCK_N", "0"))
a.impl = int(os.environ.get("IMPL", "0"))
a.max_ctas = int(os.environ.get("MAX_CTAS", "0"))
if res:
    a.res[0] = x.data_ptr()
    a.res_ctot = ctot
for _ in range(3):
    _lib.check(L.b200dn_igemm(a, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    L.b200dn_igemm(a, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
taps = 9 if kind == "conv" else 4
pix = B * H * H if kind != "down" else B * Ho * Ho
print(f"{kind} B{B} {H}x{H} cin {cin} cout {cout} ctot {ctot} {prec_name} res {res}: {ms * 1e3:.1f} us, "
      f"{2.0 * pix * taps * cin * cout / ms / 1e9:.1f} TFLOP/s", flush=True)
