set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_metrics_noise.py tests/test_gpu_sidd.py tests/test_gpu_network.py -m gpu -x -q > gpurun_out/r02_pytest_metrics.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_metrics.log
tail -30 gpurun_out/r02_pytest_metrics.log
python tools/hbm_kernels_bench.py > gpurun_out/r02_hbm_kernels.log 2>&1
cat gpurun_out/r02_hbm_kernels.log
