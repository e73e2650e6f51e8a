"""Run one igemm layer a few times (for ncu): python tools/one_layer.py B H W cin cout block_n m_tiles [prec]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import bench_layer  # noqa: E402

B, H, W, cin, cout, bn, mt = (int(v) for v in sys.argv[1:8])
prec = int(sys.argv[8]) if len(sys.argv) > 8 else 0
impl = int(sys.argv[9]) if len(sys.argv) > 9 else 0
ctot = int(sys.argv[10]) if len(sys.argv) > 10 else 0
bench_layer(B, H, W, cin, cout, prec=prec, block_n=bn, mt=mt, iters=3, impl=impl, ctot=ctot, out_ctot=ctot)
