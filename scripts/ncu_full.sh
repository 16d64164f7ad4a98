# one `ncu --set full` capture of each GEMM kernel of a cfg-3 step (source-level stall samples)
set -u
mkdir -p gpurun_out
# launches per step: prep, FWD, combine, BWDG, DW, DX, dx_finish (7); skip the warm-up steps
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 16 -c 4 \
  -f -o gpurun_out/r2_full python bench.py --no-cpu-baseline --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/r2_full.ncu-rep
tail -3 gpurun_out/ncu_full.log
