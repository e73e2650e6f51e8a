cd $GRAFT_REPO_ROOT
python tools/forward_once.py sampler 32 16 fp16 > gpurun_out/r02_sampler_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 120 --csv --log-file gpurun_out/r02_sampler_launches_raw.csv python tools/forward_once.py sampler 32 16 fp16 > gpurun_out/ncu_sampler_r02.log 2>&1
tail -2 gpurun_out/ncu_sampler_r02.log; wc -l gpurun_out/r02_sampler_launches_raw.csv
