cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/ncu_layer.py conv 128 128 32 160 32 fp16 > gpurun_out/l1_plain.log 2>&1 || exit 1
cat gpurun_out/l1_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_slab -s 3 -c 1 -o gpurun_out/l1_c128_n32 -f python tools/ncu_layer.py conv 128 128 32 160 32 fp16 > gpurun_out/ncu_l1.log 2>&1
python tools/ncu_layer.py conv 64 256 64 320 32 fp16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_slab -s 3 -c 1 -o gpurun_out/l2_c256_n64 -f python tools/ncu_layer.py conv 64 256 64 320 32 fp16 > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out/*.ncu-rep
