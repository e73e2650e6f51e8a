cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 1 --skip-diffusion --skip-cpu > gpurun_out/r02_bench_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_raw.csv python bench.py --steps 2 --warmup 1 --skip-diffusion --skip-cpu > gpurun_out/ncu_bench_r02.log 2>&1
tail -2 gpurun_out/ncu_bench_r02.log
python tools/forward_once.py sampler 32 16 fp16 > gpurun_out/r02_sampler_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2400 -c 140 --csv --log-file gpurun_out/r02_sampler_launches_raw.csv python tools/forward_once.py sampler 32 16 fp16 > gpurun_out/ncu_sampler_r02.log 2>&1
tail -2 gpurun_out/ncu_sampler_r02.log
wc -l gpurun_out/r02_bench_launches_raw.csv gpurun_out/r02_sampler_launches_raw.csv
