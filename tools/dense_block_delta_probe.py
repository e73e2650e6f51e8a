"""Diagnostic for the fused dense-block kernel: conv_3 = one delta tap on the x channels (identity PReLU, zero bias),
so out[o](y, x) - x[o](y, x) must equal x[o](y + ky - 1, x + kx - 1).  Reports, for mismatching outputs, which input
(channel, dy, dx) the kernel actually produced."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vub_image_denoising_b200 import _lib  # noqa: E402

ky, kx = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1, 1)
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (26, 18)
prec, dt, dev = _lib.PREC_FP16, torch.float16, "cuda"
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B = 1
# x[c](y, x) encodes its own coordinates exactly in fp16: c + 32 * (y * 32 + x) would overflow; use small exact ints
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
x = torch.zeros(B, H, W, 80)
for c in range(32):
    x[0, :, :, c] = ((yy * 7 + xx * 3 + c * 11) % 61).float() / 4.0        # exact in fp16, distinct-ish patterns
x = x.to(dev).to(dt)
ws = [torch.zeros(16, 32, 3, 3, device=dev), torch.zeros(16, 48, 3, 3, device=dev), torch.zeros(16, 64, 3, 3, device=dev),
      torch.zeros(32, 80, 3, 3, device=dev)]
for o in range(32):
    ws[3][o, o, ky, kx] = 1.0
bs = [torch.zeros(co, device=dev) for co in (16, 16, 16, 32)]
ss = [torch.ones(co, device=dev) for co in (16, 16, 16, 32)]
wf = torch.empty(L.b200dn_dense_block_weight_bytes(32) // 2, dtype=torch.int16, device=dev)
_lib.check(L.b200dn_pack_dense_block_weights(ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ws[3].data_ptr(), 32, prec, wf.data_ptr(), st))
out = torch.zeros((B, H, W, 32), device=dev, dtype=dt)
a = _lib.DenseBlockArgs()
a.prec, a.B, a.H, a.W, a.channels = prec, B, H, W, 32
a.in_, a.in_ctot = x.data_ptr(), 80
a.out, a.out_ctot, a.out_coff = out.data_ptr(), 32, 0
a.wfused = wf.data_ptr()
for j in range(4):
    a.bias[j], a.slope[j] = bs[j].data_ptr(), ss[j].data_ptr()
h = C.c_void_p()
_lib.check(L.b200dn_dense_block_prepare(C.byref(a), C.byref(h)), "prepare")
_lib.check(L.b200dn_igemm_launch(h, st), "launch")
torch.cuda.synchronize()
xs = x[0, :, :, :32].float().cpu()
conv = (out[0].float().cpu() - xs)                      # what the kernel's conv part produced
pad = torch.zeros(H + 2, W + 2, 32)
pad[1:-1, 1:-1] = xs
want = pad[ky:ky + H, kx:kx + W]
bad = (conv != want)
print(f"delta tap ({ky},{kx}) on {H}x{W}: {int(bad.sum())} / {bad.numel()} wrong")
if bad.any():
    # for a sample of wrong outputs, search which (c, dy, dx) matches
    big = torch.zeros(H + 8, W + 8, 32)
    big[4:-4, 4:-4] = xs
    idx = bad.nonzero()
    shown = 0
    for (y, xq, o) in idx[:: max(1, len(idx) // 24)].tolist():
        v = float(conv[y, xq, o])
        hits = []
        for c in range(32):
            for dy in range(-4, 5):
                for dx in range(-4, 5):
                    if float(big[4 + y + dy, 4 + xq + dx, c]) == v:
                        hits.append((c, dy, dx))
        print(f"  out ch {o:2d} ({y:2d},{xq:2d}): got {v:7.2f} want {float(want[y, xq, o]):7.2f}  matches {hits[:6]}")
        shown += 1
    print("  wrong per channel:", bad.sum(dim=(0, 1)).tolist())
    print("  wrong per row:", bad.sum(dim=(1, 2)).tolist())
    print("  wrong per col:", bad.sum(dim=(0, 2)).tolist())
