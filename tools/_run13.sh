cd $GRAFT_REPO_ROOT
timeout 100 python tools/dense_block_delta_probe.py 1 1 > gpurun_out/r02_dblk_delta.log 2>&1
timeout 100 python tools/dense_block_delta_probe.py 0 2 >> gpurun_out/r02_dblk_delta.log 2>&1
cat gpurun_out/r02_dblk_delta.log
