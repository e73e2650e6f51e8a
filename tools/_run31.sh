cd $GRAFT_REPO_ROOT
for i in 1 2; do
  (cd _ab/prev && python bench.py --steps 10 --warmup 3 --skip-cpu --skip-diffusion 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PREV', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks']['sm_mhz'])")
  python bench.py --steps 10 --warmup 3 --skip-cpu --skip-diffusion 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HEAD', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks']['sm_mhz'])"
done
