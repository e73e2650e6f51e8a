cd $GRAFT_REPO_ROOT
timeout 120 python tools/dense_block_probe.py 1 26 18 fp16 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 120 python tools/dense_block_probe.py 3 40 72 bf16 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 200 python tools/dense_block_probe.py 32 256 256 fp16 --bench 2>&1 | grep -v "per-channel\|got\[\|ref\["
timeout 100 python tools/dense_block_timeline.py 32 > gpurun_out/r02_dblk_timeline_v2.log 2>&1
sed -n 1,130p gpurun_out/r02_dblk_timeline_v2.log
