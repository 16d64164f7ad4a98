set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag5.txt
: > $OUT
echo "== p2p transport on one GPU" >> $OUT
timeout 900 python -m pytest tests/test_p2p_gpu.py -x -q -m gpu 2>&1 | tail -15 >> $OUT
echo "== head suite" >> $OUT
timeout 900 python -m pytest tests/test_head_gpu.py -x -q -m gpu 2>&1 | tail -5 >> $OUT
cat $OUT
