"""First-contact probe for a fresh GPU box: runs a few igemm cases with verbose diagnostics, then a
layer-shape microbenchmark.  Writes gpurun_out/probe.log-style output to stdout."""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import vub_image_denoising_b200 as b2  # noqa: E402
from vub_image_denoising_b200 import _lib  # noqa: E402
from helpers import effective_input, effective_weight, from_planes, to_planes, dt16, two_planes  # noqa: E402

DEV = "cuda"


def conv_case(B, H, W, cin, cout, prec, mode=_lib.MODE_CONV3X3, block_n=0, impl=0, mt=0):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, cin, H, W, generator=g).to(DEV)
    k = 3 if mode == _lib.MODE_CONV3X3 else 2
    w = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (k * k * cin)) ** 0.5).to(DEV)
    bias = (torch.randn(cout, generator=g) * 0.1).to(DEV)
    slope = torch.full((cout,), 0.25, device=DEV)
    x_hi, x_lo = to_planes(x, prec)
    wp = torch.ops.b200dn.pack_weight(w, prec, False)
    Ho, Wo = (H, W) if mode == _lib.MODE_CONV3X3 else (H // 2, W // 2)
    out_hi = torch.zeros((B, Ho, Wo, cout), dtype=torch.int16, device=DEV)
    out_lo = torch.zeros_like(out_hi) if two_planes(prec) else None
    torch.ops.b200dn.conv_igemm(x_hi, x_lo, wp, bias, slope, mode, prec, cin, cout, out_hi, out_lo, 0, None, None, block_n, 0, mt, impl)
    torch.cuda.synchronize()
    xe, we = effective_input(x, prec).double().cpu(), effective_weight(w, prec).double().cpu()
    ref = F.conv2d(xe, we, bias.double().cpu(), padding=1) if mode == _lib.MODE_CONV3X3 else F.conv2d(xe, we, bias.double().cpu(), stride=2)
    ref = F.prelu(ref, slope.double().cpu())
    got = from_planes(out_hi, out_lo, prec, 0, cout).double().cpu()
    err = (got - ref).abs()
    rel = float(err.max() / ref.abs().max())
    print(f"  mode {mode} impl {impl} mt {mt} prec {prec} B{B} {H}x{W} cin {cin} cout {cout} bn {block_n}: max err {float(err.max()):.3e} (rel {rel:.2e}), "
          f"ref absmax {float(ref.abs().max()):.3f}, got absmax {float(got.abs().max()):.3f}", flush=True)
    if rel > 0.02:
        # locate the damage: per-channel and per-position error maps
        e_c = err.amax(dim=(0, 2, 3))
        print("   per-channel max err (first 16):", [f"{v:.2e}" for v in e_c[:16].tolist()])
        e_p = err.amax(dim=(0, 1))
        print("   rows with err:", (e_p.amax(dim=1) > 0.05).nonzero().flatten()[:20].tolist())
        print("   cols with err:", (e_p.amax(dim=0) > 0.05).nonzero().flatten()[:20].tolist())
        print("   got[0,:4,0,:4]:", got[0, :4, 0, :4].tolist())
        print("   ref[0,:4,0,:4]:", ref[0, :4, 0, :4].tolist())
    return rel


def bench_layer(B, H, W, cin, cout, prec=_lib.PREC_BF16, mode=_lib.MODE_CONV3X3, block_n=0, iters=10, res=False, mt=0, impl=0,
                ctot=0, out_ctot=0):
    d = dt16(prec)
    ctot = ctot or cin
    out_ctot = out_ctot or cout
    x_hi = (torch.randn(B, H, W, ctot, device=DEV) * 0.5).to(d).view(torch.int16)
    x_lo = torch.zeros_like(x_hi) if two_planes(prec) else None
    k = 3 if mode == _lib.MODE_CONV3X3 else 2
    w = torch.randn(cout, cin, k, k, device=DEV) * 0.02
    wp = torch.ops.b200dn.pack_weight(w, prec, False)
    bias = torch.zeros(cout, device=DEV)
    slope = torch.full((cout,), 0.25, device=DEV)
    Ho, Wo = (H, W) if mode == _lib.MODE_CONV3X3 else (H // 2, W // 2)
    out_hi = torch.empty((B, Ho, Wo, out_ctot), dtype=torch.int16, device=DEV)
    out_lo = torch.empty_like(out_hi) if two_planes(prec) else None
    r_hi = x_hi if res else None
    a = _lib.IgemmArgs()
    a.mode, a.prec, a.B, a.H, a.W, a.cin, a.cout = mode, prec, B, H, W, cin, cout
    a.in_[0], a.in_[1] = x_hi.data_ptr(), (x_lo.data_ptr() if x_lo is not None else None)
    a.in_ctot = ctot
    a.wpacked, a.bias, a.slope = wp.data_ptr(), bias.data_ptr(), slope.data_ptr()
    a.out_kind = 0
    a.out[0], a.out[1] = out_hi.data_ptr(), (out_lo.data_ptr() if out_lo is not None else None)
    a.out_ctot, a.out_coff = out_ctot, 0
    if res:
        a.res[0] = r_hi.data_ptr()
        a.res_ctot = ctot
    a.block_n = block_n
    a.m_tiles = mt
    a.impl = impl
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(L.b200dn_igemm(a, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.b200dn_igemm(a, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    taps = 9 if mode == _lib.MODE_CONV3X3 else 4
    flops = 2.0 * B * Ho * Wo * taps * cin * cout
    n_mma = {0: 1, 1: 1, 2: 2, 3: 3, 4: 2}[prec]
    print(f"  bench B{B} {H}x{W} cin {cin:4d} cout {cout:4d} bn {block_n:3d} mt {mt} impl {impl} prec {prec}: {ms:8.3f} ms  "
          f"{flops / ms / 1e9:8.1f} TFLOP/s (useful)  {flops * n_mma / ms / 1e9:8.1f} TFLOP/s (issued)", flush=True)
    return ms


def main():
    print(torch.cuda.get_device_name(0), "SMs", _lib.lib().b200dn_sm_count(), flush=True)
    print("== correctness probes", flush=True)
    worst = 0.0
    worst = max(worst, conv_case(1, 8, 16, 64, 16, _lib.PREC_BF16))
    worst = max(worst, conv_case(1, 8, 16, 16, 16, _lib.PREC_BF16))
    worst = max(worst, conv_case(1, 16, 32, 128, 64, _lib.PREC_BF16))
    worst = max(worst, conv_case(2, 24, 40, 80, 48, _lib.PREC_BF16))
    worst = max(worst, conv_case(1, 16, 32, 64, 256, _lib.PREC_FP16))
    worst = max(worst, conv_case(1, 16, 32, 64, 64, _lib.PREC_BF16X2))
    worst = max(worst, conv_case(1, 16, 32, 64, 64, _lib.PREC_BF16X3))
    worst = max(worst, conv_case(1, 16, 32, 64, 128, _lib.PREC_BF16, mode=_lib.MODE_DOWN2X2))
    print(f"worst relative error {worst:.3e}", flush=True)
    if worst > 0.02:
        print("PROBE FAILED: skipping the benchmark", flush=True)
        return 1
    print("== layer microbenchmarks (RDUNet(128) shapes, B=8)", flush=True)
    for cin, cout in ((128, 64), (192, 64), (256, 64), (320, 128)):
        bench_layer(8, 256, 256, cin, cout, mt=1)
        bench_layer(8, 256, 256, cin, cout, mt=2)
    for cin, cout in ((256, 128), (640, 256)):
        bench_layer(8, 128, 128, cin, cout, mt=1)
    bench_layer(8, 128, 128, 256, 128, mt=2)
    for cin, cout in ((512, 256), (1280, 512)):
        bench_layer(8, 64, 64, cin, cout)
    for cin, cout in ((1024, 512), (2560, 1024)):
        bench_layer(8, 32, 32, cin, cout)
    print("== N-tile sweep at 128x128, cin 256", flush=True)
    for bn in (32, 64, 128, 256):
        bench_layer(8, 128, 128, 256, 256, block_n=bn)
    print("== RDUNet(32) shapes, B=8", flush=True)
    for cin, cout in ((32, 16), (80, 32)):
        bench_layer(8, 256, 256, cin, cout, mt=1)
        bench_layer(8, 256, 256, cin, cout, mt=2)
    bench_layer(8, 32, 32, 640, 256)
    return 0


if __name__ == "__main__":
    sys.exit(main())
