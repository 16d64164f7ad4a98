set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag12.txt
: > $OUT
export ASM_B200_LIB=$PWD/tf_face_toolbox_b200/lib/bringup/libasoftmax_b200.so
for d in 0 4; do
  ASM_UMMA_DEBUG=$d timeout 200 python scripts/kernel_times.py 512 512 85742 40 2>&1 | tail -1 | cut -c28-220 >> $OUT
done
for d in 0 4; do
ASM_UMMA_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,lts__t_bytes.sum,sm__cycles_active.avg,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:umma_kernel -s 8 -c 1 --csv --log-file gpurun_out/diag12_ncu_$d.csv python scripts/kernel_times.py 512 512 85742 3 > /dev/null 2>&1
grep -E "umma_kernel" gpurun_out/diag12_ncu_$d.csv | awk -F'","' '{print $5, $(NF-3), $(NF-2), $(NF)}' | cut -c1-200 >> $OUT
done
cat $OUT
