# round-2: validate the DW TMA-store path / traversal order / L2 hints, then A/B them on one box
set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag2.txt
: > $OUT
echo "== gpu test suite" >> $OUT
timeout 900 python -m pytest tests/test_head_gpu.py -x -q -m gpu 2>&1 | tail -8 >> $OUT
echo "== p2p transport on one GPU" >> $OUT
timeout 600 python -m pytest tests/test_p2p_gpu.py -x -q -m gpu 2>&1 | tail -8 >> $OUT
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --no-cpu-baseline --steps 50 --warmup 5 2>gpurun_out/diag2_$name.err | tail -1 > gpurun_out/diag2_$name.json
  python - <<PY >> $OUT 2>&1
import json
try:
    d=json.loads(open("gpurun_out/diag2_$name.json").read())
    print("$name", round(d["value"]), round(d["ms_per_step"]*1000,1), "us e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], {k["kernel"][:8]:round(k["ms"]*1000,1) for k in d["kernels"]})
except Exception as e:
    print("$name FAILED", e)
PY
}
run default X=0
run no_order ASM_L2_ORDER=0
run no_hints ASM_L2_HINTS=0
run no_dwtma ASM_DW_TMA=0
run all_off ASM_L2_ORDER=0 ASM_L2_HINTS=0 ASM_DW_TMA=0
run default2 X=0
M=dram__bytes_read.sum,dram__bytes_write.sum
for v in on off; do
  if [ $v = off ]; then export ASM_L2_ORDER=0 ASM_L2_HINTS=0 ASM_DW_TMA=0; fi
  timeout 300 ncu --metrics $M --cache-control none --clock-control none -s 70 -c 14 --csv --log-file gpurun_out/diag2_ncu_$v.csv \
    python bench.py --no-cpu-baseline --no-graph --steps 2 --warmup 3 > gpurun_out/diag2_ncu_$v.log 2>&1
done
cat $OUT
