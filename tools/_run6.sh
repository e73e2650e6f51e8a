set -x
cd $GRAFT_REPO_ROOT
python tools/ncu_metric_kernels.py > gpurun_out/r02_metric_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ssim_kernel -s 2 -c 1 -o gpurun_out/r02_ssim python tools/ncu_metric_kernels.py > gpurun_out/ncu_ssim.log 2>&1
tail -3 gpurun_out/ncu_ssim.log
