set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag8.txt
: > $OUT
echo "== head + p2p tests" >> $OUT
timeout 1200 python -m pytest tests/test_head_gpu.py tests/test_p2p_gpu.py -x -q -m gpu 2>&1 | tail -8 >> $OUT
for i in 1 2; do
timeout 200 python scripts/kernel_times.py 512 512 85742 40 2>&1 | tail -1 >> $OUT
done
timeout 200 python scripts/kernel_times.py 512 512 10752 40 2>&1 | tail -1 >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 16 -c 4 \
  -f -o gpurun_out/r2_full_b python bench.py --no-cpu-baseline --no-graph --no-cfg4 --steps 2 --warmup 3 > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out/r2_full_b.ncu-rep >> $OUT
cat $OUT
