# E1: is the GEMM mainloop bound by DRAM or by the L2 -> SM fabric?  Mainloop-only (bring-up build,
# ASM_UMMA_DEBUG=1) and full kernels at class counts that give whole rounds of 74 CTA pairs.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag4.txt
: > $OUT
export ASM_B200_LIB=$PWD/tf_face_toolbox_b200/lib/bringup/libasoftmax_b200.so
for C in 18944 37888 75776 85742 151552; do
  for dbg in 1 0; do
    ASM_UMMA_DEBUG=$dbg timeout 120 python scripts/kernel_times.py 512 512 $C 30 2>&1 | tail -1 >> $OUT
  done
done
for C in 18944 85742; do
  ASM_UMMA_DEBUG=5 timeout 120 python scripts/kernel_times.py 512 512 $C 30 2>&1 | tail -1 >> $OUT
done
ASM_UMMA_DEBUG=1 timeout 120 python scripts/kernel_times.py 2048 512 85742 20 2>&1 | tail -1 >> $OUT
ASM_UMMA_DEBUG=0 timeout 120 python scripts/kernel_times.py 2048 512 85742 20 2>&1 | tail -1 >> $OUT
cat $OUT
