mkdir -p gpurun_out
( for v in "ASM_DW_TMA=1" "ASM_DW_TMA=0"; do echo "$v"; env $v ASM_PREP_STREAM=0 timeout 300 python scripts/diag_bf16.py; done
  for v in "ASM_PREP_STREAM=0" "ASM_PREP_STREAM=1" "ASM_PREP_STREAM=0 ASM_UMMA_BN=128"; do env $v timeout 300 python scripts/fwd_times.py; done ) > gpurun_out/diag_bf16.txt 2>&1
