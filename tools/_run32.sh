cd $GRAFT_REPO_ROOT
for L in "conv 128 64 32 160 32 fp16" "conv 128 128 32 160 32 fp16" "conv 128 160 64 160 32 fp16 1 160 0" "conv 64 128 64 320 32 fp16" "conv 64 320 128 320 32 fp16 1 320 0"; do
  echo "== $L"
  echo -n "default      "; python tools/ncu_layer.py $L
  echo -n "mt=1         "; M_TILES=1 python tools/ncu_layer.py $L
  echo -n "wres off     "; B200DN_SLAB_WRES=0 python tools/ncu_layer.py $L
  echo -n "pairs impl=3 "; IMPL=3 python tools/ncu_layer.py $L
  echo -n "pairs mt=1   "; IMPL=3 M_TILES=1 python tools/ncu_layer.py $L
  echo -n "pitch 16     "; B200DN_SLAB_PITCH=16 python tools/ncu_layer.py $L
done
