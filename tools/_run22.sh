cd $GRAFT_REPO_ROOT
timeout 120 python tools/dense_block_probe.py 3 40 72 bf16 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 200 python tools/dense_block_probe.py 32 256 256 fp16 --bench 2>&1 | grep -v "per-channel\|got\[\|ref\["
