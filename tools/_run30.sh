cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_layers.py tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -3
python tools/layer_times.py 32 32 fp16 > gpurun_out/r02_layer_times_f32_warparrive.log 2>&1
tail -12 gpurun_out/r02_layer_times_f32_warparrive.log
sed -n 1,14p gpurun_out/r02_layer_times_f32_warparrive.log
python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print(d['diffusion']['value'], d['diffusion']['batch1'])
PY
