set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag10.txt
: > $OUT
for p in 0 30 37 42 46 52 0; do
  ASM_DW_PAIRS=$p timeout 200 python scripts/kernel_times.py 512 512 85742 60 2>&1 | tail -1 | cut -c1-80 >> $OUT
done
ASM_DW_PAIRS=42 timeout 300 python scripts/try_head.py 512 512 85742 bf16 2>&1 | tail -1 >> $OUT
for p in 0 42; do
  ASM_DW_PAIRS=$p timeout 200 python scripts/kernel_times.py 2048 512 85742 20 2>&1 | tail -1 | cut -c1-80 >> $OUT
done
cat $OUT
