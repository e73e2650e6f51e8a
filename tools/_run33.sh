cd $GRAFT_REPO_ROOT
for L in "conv 256 32 32 32 32 fp16 0 32 0" "conv 256 96 32 96 32 fp16 0 32 0" "conv 128 96 32 160 32 fp16" "conv 256 32 32 32 2 fp16 0 32 0" "conv 128 64 32 160 2 fp16" "conv 128 128 32 160 4 fp16"; do
  echo "== $L"
  echo -n "default      "; python tools/ncu_layer.py $L
  echo -n "pairs impl=3 "; IMPL=3 python tools/ncu_layer.py $L
done
