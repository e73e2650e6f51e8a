set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
L1="python tools/ncu_layer.py conv 256 32 16 80 32 fp16"
L2="python tools/ncu_layer.py conv 256 80 32 80 32 fp16 1 80 0"
L3="python tools/ncu_layer.py up 128 64 64 32 fp16"
L4="python tools/ncu_layer.py conv 128 64 32 160 32 fp16"
$L1 > gpurun_out/r02_l1_plain.log 2>&1 && $L2 > gpurun_out/r02_l2_plain.log 2>&1 && $L3 > gpurun_out/r02_l3_plain.log 2>&1 && $L4 > gpurun_out/r02_l4_plain.log 2>&1 || exit 1
cat gpurun_out/r02_l?_plain.log
python tools/layer_times.py 32 32 fp16 > gpurun_out/r02_base_layer_times_f32.log 2>&1
for b in 1 2 16; do python tools/sampler_latency.py $b >> gpurun_out/r02_base_sampler_latency.log 2>&1; done
cat gpurun_out/r02_base_sampler_latency.log
ncu --set full --clock-control none --import-source on -k regex:conv3x3_slab -s 4 -c 1 -o gpurun_out/r02_base_conv_32_16 $L1 > gpurun_out/ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_slab -s 4 -c 1 -o gpurun_out/r02_base_conv_80_32 $L2 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 4 -c 1 -o gpurun_out/r02_base_up_64 $L3 > gpurun_out/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_slab -s 4 -c 1 -o gpurun_out/r02_base_conv_64_32_l1 $L4 > gpurun_out/ncu_l4.log 2>&1
tail -3 gpurun_out/ncu_l?.log
ls -la gpurun_out
