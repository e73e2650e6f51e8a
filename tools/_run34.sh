cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_metrics_noise.py tests/test_gpu_network.py tests/test_gpu_layers.py -m gpu -x -q 2>&1 | tail -3
python tools/hbm_kernels_bench.py 2>&1 | grep -i "welch\|ssim\|psnr"
python tools/layer_times.py 32 32 fp16 2>&1 | tail -11
