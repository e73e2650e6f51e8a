cd $GRAFT_REPO_ROOT
timeout 100 python tools/dense_block_timeline.py 32 > gpurun_out/r02_dblk_timeline_v3.log 2>&1
sed -n 1,150p gpurun_out/r02_dblk_timeline_v3.log
