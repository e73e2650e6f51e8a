"""CTA-pair (cta_group::2) slab kernel: correctness probes, then impl 2 (one CTA) vs impl 3 (CTA pair) per layer shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import bench_layer, conv_case  # noqa: E402
from vub_image_denoising_b200 import _lib  # noqa: E402


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    worst = 0.0
    for (B, H, W, cin, cout, prec, mt) in [(1, 16, 32, 64, 64, _lib.PREC_BF16, 1), (2, 32, 32, 128, 64, _lib.PREC_BF16, 2),
                                           (3, 24, 40, 160, 128, _lib.PREC_BF16, 1), (1, 16, 16, 64, 512, _lib.PREC_FP16, 1),
                                           (2, 64, 64, 320, 128, _lib.PREC_BF16, 2), (1, 32, 32, 96, 32, _lib.PREC_BF16X2, 0),
                                           (1, 8, 8, 128, 256, _lib.PREC_BF16X3, 0)]:
        worst = max(worst, conv_case(B, H, W, cin, cout, prec, impl=3, mt=mt))
    print(f"worst relative error {worst:.3e}", flush=True)
    if worst > 0.02:
        print("PROBE FAILED", flush=True)
        return 1
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    print(f"== RDUNet(128) layer shapes, B={B}: impl 2 (one CTA per tile) vs impl 3 (CTA pair)", flush=True)
    for (hw, cin, cout) in [(256, 128, 64), (256, 192, 64), (256, 256, 64), (256, 320, 128), (256, 384, 128),
                            (128, 256, 128), (128, 384, 128), (128, 512, 128), (128, 640, 256), (128, 768, 256),
                            (64, 512, 256), (64, 1280, 512), (32, 1024, 512), (32, 2560, 1024)]:
        for impl in (2, 3):
            bench_layer(B, hw, hw, cin, cout, impl=impl, iters=5)
    return 0


if __name__ == "__main__":
    sys.exit(main())
