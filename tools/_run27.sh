cd $GRAFT_REPO_ROOT
# level-0 transposed conv of RDUNet(32): 64 -> 4 x 64, written at channel offset 32 of a 96-channel buffer
L="python tools/ncu_layer.py up 128 64 64 32 fp16 0 96 32"
echo "default"; $L
echo "staged always"; B200DN_EPI_STAGED=1 $L
echo "mt=1"; M_TILES=1 $L
echo "mt=1 staged off"; M_TILES=1 B200DN_EPI_STAGED=0 $L
echo "level 1: 128 -> 4x128 @64 coff 64 of 192"; python tools/ncu_layer.py up 64 128 128 32 fp16 0 192 64
echo " mt=1"; M_TILES=1 python tools/ncu_layer.py up 64 128 128 32 fp16 0 192 64
echo " staged always"; B200DN_EPI_STAGED=1 python tools/ncu_layer.py up 64 128 128 32 fp16 0 192 64
