cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_layers.py tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -3
python tools/hbm_kernels_bench.py 2>&1 | grep -i "conv_in"
for b in 16; do python tools/sampler_latency.py $b; done
