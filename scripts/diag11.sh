set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag11.txt
: > $OUT
ASM_BW_MODE=2 timeout 120 python scripts/try_head.py 512 512 85742 bf16 2>&1 | tail -1 >> $OUT
for cfgs in "1 0,0" "2 26,24" "2 30,22" "2 24,26" "2 28,26" "2 32,22" "2 22,28"; do
  set -- $cfgs
  ASM_BW_MODE=$1 ASM_BW_SPLIT=$2 timeout 200 python scripts/kernel_times.py 512 512 85742 40 2>&1 | tail -1 | cut -c28-200 >> $OUT
done
timeout 900 python -m pytest tests/test_head_gpu.py tests/test_p2p_gpu.py -x -q -m gpu 2>&1 | tail -4 >> $OUT
cat $OUT
