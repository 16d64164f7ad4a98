# which call of test_bf16_embeddings_at_the_boundary has the wrong dW, and where (the clipped-TMA-box bug)
mkdir -p gpurun_out
( for v in "ASM_DW_TMA=1" "ASM_DW_TMA=0"; do echo "$v"; env $v timeout 300 python scripts/diag_bf16.py; done ) > gpurun_out/diag_bf16.txt 2>&1
