cd $GRAFT_REPO_ROOT
timeout 100 python tools/dense_block_delta_probe.py 1 1 2>&1 | head -8
timeout 100 python tools/dense_block_delta_probe.py 0 2 2>&1 | head -8
timeout 100 python tools/dense_block_delta_probe.py 2 0 64 64 2>&1 | head -8
for m in x o0 o1 o2; do echo "== mask $m"; timeout 120 python tools/dense_block_probe.py 1 64 64 fp16 --mask=$m 2>&1 | grep -v "per-channel\|got\[\|ref\["; done
timeout 120 python tools/dense_block_probe.py 1 26 18 fp16 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 120 python tools/dense_block_probe.py 2 64 64 fp16 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 120 python tools/dense_block_probe.py 3 40 72 bf16 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 200 python tools/dense_block_probe.py 32 256 256 fp16 --bench 2>&1 | grep -v "per-channel\|got\[\|ref\["
