set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
python tools/forward_latency.py 128 1 > gpurun_out/r02_forward_latency.log 2>&1
python tools/forward_latency.py 32 1 >> gpurun_out/r02_forward_latency.log 2>&1
cat gpurun_out/r02_forward_latency.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
cat gpurun_out/r02_bench_a.json
tail -5 gpurun_out/r02_bench_a.err
python tools/layer_times.py 32 32 fp16 > gpurun_out/r02_layer_times_f32_satflag.log 2>&1
tail -12 gpurun_out/r02_layer_times_f32_satflag.log
