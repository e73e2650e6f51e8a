export B200DN_LIB=$PWD/vub_image_denoising_b200/libb200dn_diag.so
export B200DN_GRAPH=0
for stop in 1 2; do
  echo "== pair, stop stage $stop"
  B200DN_DENSE_STOP=$stop B200DN_DENSE_PAIR=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^dense_block_kernel -c 3 python tools/forward_once.py rdunet 32 2 fp16 2>&1 | grep -v "^==PROF== Prof\|WARNING" | grep -E "duration|ERROR|not profiled|ok" | head
done
