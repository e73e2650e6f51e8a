cd $GRAFT_REPO_ROOT
for m in x o0 o1 o2; do
  echo "== mask $m" >> gpurun_out/r02_dblk_probe2.log
  timeout 120 python tools/dense_block_probe.py 1 26 18 fp16 --mask=$m >> gpurun_out/r02_dblk_probe2.log 2>&1
  timeout 120 python tools/dense_block_probe.py 1 64 64 fp16 --mask=$m >> gpurun_out/r02_dblk_probe2.log 2>&1
done
grep -v "per-channel\|got\[\|ref\[" gpurun_out/r02_dblk_probe2.log
