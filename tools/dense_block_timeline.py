"""Timeline of CTA 0 of the fused dense-block kernel (issuer and epilogue hand-offs, SM clocks):
   python tools/dense_block_timeline.py [B=32]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vub_image_denoising_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = W = 256
prec, dt, dev = _lib.PREC_FP16, torch.float16, "cuda"
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
x = (torch.randn(B, H, W, 80, device=dev) * 0.5).to(dt)
ws = [torch.randn(co, ci, 3, 3, device=dev) * 0.05 for co, ci in ((16, 32), (16, 48), (16, 64), (32, 80))]
bs = [torch.zeros(co, device=dev) for co in (16, 16, 16, 32)]
ss = [torch.full((co,), 0.25, device=dev) for co in (16, 16, 16, 32)]
wf = torch.empty(L.b200dn_dense_block_weight_bytes(32) // 2, dtype=torch.int16, device=dev)
_lib.check(L.b200dn_pack_dense_block_weights(ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ws[3].data_ptr(), 32, prec, wf.data_ptr(), st))
out = torch.zeros((B, H, W, 32), device=dev, dtype=dt)
tl = torch.zeros(3 * 2730, dtype=torch.int64, device=dev)
a = _lib.DenseBlockArgs()
a.prec, a.B, a.H, a.W, a.channels = prec, B, H, W, 32
a.in_, a.in_ctot = x.data_ptr(), 80
a.out, a.out_ctot, a.out_coff = out.data_ptr(), 32, 0
a.wfused = wf.data_ptr()
for j in range(4):
    a.bias[j], a.slope[j] = bs[j].data_ptr(), ss[j].data_ptr()
h0 = C.c_void_p()
_lib.check(L.b200dn_dense_block_prepare(C.byref(a), C.byref(h0)))
for _ in range(2):
    _lib.check(L.b200dn_igemm_launch(h0, st))
torch.cuda.synchronize()
a.timeline = tl.data_ptr()
h = C.c_void_p()
_lib.check(L.b200dn_dense_block_prepare(C.byref(a), C.byref(h)))
_lib.check(L.b200dn_igemm_launch(h, st))
torch.cuda.synchronize()
t = tl.cpu().tolist()
ev = []
for r in range(3):
    base = r * 2730
    for it in range(16):
        for ps in range(4):
            for tile in range(6):
                for kind in range(4):
                    clk = t[base + 2 + it * 160 + ps * 40 + tile * 4 + kind]
                    if clk:
                        ev.append((clk, (r + 1) * 100000 + it * 1000 + ps * 100 + tile * 10 + (9 if kind == 3 else kind)))
ev.sort()
n = len(ev)
t0 = ev[0][0]
names = {1: "issuer", 2: "epi0  ", 3: "epi1  "}
kinds = {1: {0: "deps ok", 1: "issued", 9: "X full"}, 2: {0: "tfull", 1: "ldtm done", 2: "arrived"}, 3: {0: "tfull", 1: "ldtm done", 2: "arrived"}}
print(f"{n} events; regions of CTA 0; clocks relative to the first event")
for clk, e in ev:
    role, rest = divmod(e, 100000)
    it, rest = divmod(rest, 1000)
    ps, rest = divmod(rest, 100)
    tile, kind = divmod(rest, 10)
    if it in (3, 4):
        print(f"{clk - t0:8d}  {names[role]} region {it} pass {ps} tile {tile}  {kinds[role].get(kind, kind)}")
