cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_layers.py tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -3
python tools/layer_times.py 32 32 fp16 > gpurun_out/r02_layer_times_f32_walker.log 2>&1
cat gpurun_out/r02_layer_times_f32_walker.log | tail -75
