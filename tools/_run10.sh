set -x
cd $GRAFT_REPO_ROOT
timeout 120 python tools/dense_block_probe.py 1 26 18 fp16 > gpurun_out/r02_dblk_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dblk_probe.log
timeout 120 python tools/dense_block_probe.py 2 64 64 fp16 >> gpurun_out/r02_dblk_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dblk_probe.log
timeout 120 python tools/dense_block_probe.py 3 40 72 bf16 >> gpurun_out/r02_dblk_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dblk_probe.log
timeout 200 python tools/dense_block_probe.py 32 256 256 fp16 --bench >> gpurun_out/r02_dblk_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dblk_probe.log
cat gpurun_out/r02_dblk_probe.log
