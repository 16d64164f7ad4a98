set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag3.txt
: > $OUT
for v in 2 3 1; do
  echo "== ASM_DW_TMA=$v (2: even rows only, 3: odd rows only; results are partial by design)" >> $OUT
  ASM_DW_TMA=$v timeout 120 python scripts/try_head.py 300 192 778 bf16 2>&1 | tail -2 >> $OUT
  ASM_DW_TMA=$v timeout 120 python scripts/try_head.py 256 128 1000 bf16 2>&1 | tail -2 >> $OUT
done
cat $OUT
