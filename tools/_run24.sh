cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu3.log
tail -5 gpurun_out/r02_pytest_gpu3.log
for b in 1 2 16; do python tools/sampler_latency.py $b; done 2>&1 | tee gpurun_out/r02_sampler_latency_wprefetch.log
python tools/forward_latency.py 128 1 2>&1 | tee -a gpurun_out/r02_sampler_latency_wprefetch.log
python tools/layer_times.py 32 2 fp16 > gpurun_out/r02_layer_times_f32_b2.log 2>&1; tail -14 gpurun_out/r02_layer_times_f32_b2.log
