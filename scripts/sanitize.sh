# compute-sanitizer memcheck over a few small parity tests (bf16 pair kernels, fp32 planes, C % 4 == 2, p2p on one GPU)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sanitize.txt
: > $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_head_gpu.py -x -q -m gpu \
  -k "test_bf16_shapes or test_dw_rows_that_end_off or test_edge_shapes_on_the_tensor_core_kernels or test_fused_optimizer_matches_unfused_update or test_gradient_transform" 2>&1 | tail -30 >> $OUT
echo "exit: $?" >> $OUT
