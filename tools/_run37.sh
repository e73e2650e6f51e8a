cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_layers.py -m gpu -x -q -k conv_in 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/hbm_kernels_bench.py 2>&1 | grep -i "conv_in"
for b in 1 16; do python tools/sampler_latency.py $b; done
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-diffusion 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HEAD', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks']['sm_mhz'], d['quality'])"
