cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu2.log
tail -12 gpurun_out/r02_pytest_gpu2.log
python tools/layer_times.py 32 32 fp16 > gpurun_out/r02_layer_times_f32_fused.log 2>&1
grep -E "denseblk|total|N=" gpurun_out/r02_layer_times_f32_fused.log | head -30
for b in 1 2 16; do python tools/sampler_latency.py $b; done 2>&1 | tee gpurun_out/r02_sampler_latency_fused.log
