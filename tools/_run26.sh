cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/layer_times.py 32 32 fp16 > gpurun_out/r02_layer_times_f32_walker2.log 2>&1
tail -12 gpurun_out/r02_layer_times_f32_walker2.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print(json.dumps(d['diffusion'])[:1500])
PY
