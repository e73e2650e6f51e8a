cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_network.py -m gpu -x -q 2>&1 | tail -3
for s in 64 32 16; do
  echo "== B200DN_SPLIT_N=$s"
  B200DN_SPLIT_N=$s python tools/sampler_latency.py 1
  B200DN_SPLIT_N=$s python tools/sampler_latency.py 2
done
B200DN_SPLIT_N=32 python tools/layer_times.py 32 2 fp16 2>&1 | tail -75 > gpurun_out/layer_times_b2_split32.txt
