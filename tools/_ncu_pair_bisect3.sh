echo "== probe (fused block alone + the four per-layer launches), pair, under ncu profiling ONLY the dense kernel"
B200DN_DENSE_PAIR=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^dense_block_kernel -c 6 python tools/dense_block_probe.py 2 64 64 fp16 2>&1 | grep -E "duration|ERROR|max|ok|fused|agree" | head -14
echo "== same, --bench (many launches)"
B200DN_DENSE_PAIR=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^dense_block_kernel -c 6 python tools/dense_block_probe.py 2 64 64 fp16 --bench 2>&1 | grep -E "duration|ERROR|max|ok|fused|agree" | head -14
echo "== forward, pair, profile ALL kernels (no filter), first 40"
B200DN_GRAPH=0 B200DN_DENSE_PAIR=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 python tools/forward_once.py rdunet 32 2 fp16 2>&1 | grep -E "^  [a-z].*\(|duration|ERROR" | head -60 | cut -c1-120
