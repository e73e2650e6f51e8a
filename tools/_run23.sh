cd $GRAFT_REPO_ROOT
timeout 200 python tools/dense_block_probe.py 32 256 256 fp16 --bench > gpurun_out/r02_dblk_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_block_kernel -s 2 -c 1 -o gpurun_out/r02_dblk_v3 python tools/dense_block_probe.py 32 256 256 fp16 --bench > gpurun_out/ncu_dblk3.log 2>&1
tail -2 gpurun_out/ncu_dblk3.log
