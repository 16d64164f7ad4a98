set -u
mkdir -p gpurun_out
OUT=gpurun_out/sanitize2.txt
: > $OUT
for tool in synccheck initcheck; do
  echo "== $tool" >> $OUT
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 12 \
    python -m pytest tests/test_head_gpu.py -x -q -m gpu -k "test_bf16_shapes or test_dw_rows_that_end_off" 2>&1 | grep -v "^=========     \(at\|in\|by\) \|Host Frame\|^$" | tail -40 >> $OUT
done
