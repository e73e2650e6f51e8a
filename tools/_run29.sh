cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --skip-diffusion > gpurun_out/n2_a.out 2> gpurun_out/n2_a.err; echo "rc=$?"; wc -c gpurun_out/n2_a.out; tail -3 gpurun_out/n2_a.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/n2_b.out 2> gpurun_out/n2_b.err; echo "rc=$?"; wc -c gpurun_out/n2_b.out; tail -20 gpurun_out/n2_b.err
