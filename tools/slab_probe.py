"""Correctness + speed probe of the haloed-slab conv3x3 kernel (impl=2) against the per-tap kernel (impl=1)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import conv_case, bench_layer  # noqa: E402
from vub_image_denoising_b200 import _lib  # noqa: E402

# (the descriptor base-offset experiment, profiles/r01_umma_base_offset_experiment.txt, used a B200DN_SLAB_BO switch that
# has since been removed: base_offset stays 0)
worst = 0.0
print("B200DN_SLAB_WRES =", os.environ.get("B200DN_SLAB_WRES", "(default 1)"), flush=True)
for (B, H, W, cin, cout, prec, mt) in [(1, 16, 8, 64, 16, 0, 1), (2, 32, 24, 32, 16, 1, 2), (1, 16, 16, 80, 32, 4, 1), (1, 32, 8, 96, 32, 3, 2), (1, 16, 16, 64, 64, 0, 1), (2, 24, 40, 80, 48, 0, 1),
                                       (1, 32, 32, 128, 64, 0, 2), (2, 40, 24, 160, 128, 0, 2), (1, 16, 32, 64, 256, 1, 1),
                                       (1, 32, 16, 64, 64, 2, 2), (1, 16, 16, 96, 512, 3, 1)]:
    worst = max(worst, conv_case(B, H, W, cin, cout, prec, impl=2, mt=mt))
print(f"worst relative error {worst:.3e}", flush=True)
if worst > 0.02:
    print("SLAB PROBE FAILED", flush=True)
    sys.exit(1)
if "--bench" in sys.argv:
    print("== tap (impl 1) vs slab (impl 2), RDUNet(128) shapes, B=8", flush=True)
    for (H, cin, cout) in [(256, 128, 64), (256, 256, 64), (256, 320, 128), (128, 256, 128), (128, 640, 256), (64, 1280, 512),
                           (32, 2560, 1024)]:
        for impl in (1, 2):
            bench_layer(8, H, H, cin, cout, impl=impl)
    print("== RDUNet(32) shapes, B=32, fp16, network channel strides", flush=True)
    for (H, cin, cout, ctot) in [(256, 32, 16, 80), (256, 48, 16, 80), (256, 64, 16, 80), (256, 80, 32, 80), (256, 96, 32, 96),
                                 (128, 64, 32, 160), (128, 160, 64, 160), (64, 320, 128, 320), (32, 640, 256, 640)]:
        bench_layer(32, H, H, cin, cout, impl=2, prec=_lib.PREC_FP16, ctot=ctot)
    print("== fp16x2 RDUNet(32) shapes, B=8", flush=True)
    for (H, cin, cout) in [(256, 32, 16), (256, 80, 32), (128, 160, 64)]:
        for impl in (1, 2):
            bench_layer(8, H, H, cin, cout, impl=impl, prec=_lib.PREC_FP16X2)
