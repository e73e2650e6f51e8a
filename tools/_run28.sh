cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests/test_gpu_baseline_size.py -m gpu -x -q -k "nccl or two_devices" 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print(json.dumps({k:v for k,v in d['diffusion'].items() if k in ('value','strong','batch1','e2e')}))
PY
