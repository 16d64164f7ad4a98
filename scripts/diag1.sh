# round-2 diagnostic: per-kernel times under the existing bring-up knobs (same box A/B)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag1.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv >> $OUT 2>&1
run() {  # name, env assignments...
  name=$1; shift
  env "$@" timeout 150 python bench.py --no-cpu-baseline --steps 50 --warmup 5 2>gpurun_out/diag1_$name.err | tail -1 > gpurun_out/diag1_$name.json
  python - <<PY >> $OUT 2>&1
import json
try:
    d=json.loads(open("gpurun_out/diag1_$name.json").read())
    print("$name", round(d["value"]), round(d["ms_per_step"]*1000,1), "us e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"], {k["kernel"][:8]:round(k["ms"]*1000,1) for k in d["kernels"]})
except Exception as e:
    print("$name FAILED", e)
PY
}
run base ASM_UMMA_DEBUG=0
run mainloop ASM_UMMA_DEBUG=1
run fwdr ASM_UMMA_DEBUG=4
run fwdr_mainloop ASM_UMMA_DEBUG=5
run cg1 ASM_UMMA_CG=0
run bn128 ASM_UMMA_BN=128
run nooverlap ASM_NO_OVERLAP=1
echo "== narrow tiles parity" >> $OUT
ASM_UMMA_BN=128 timeout 300 python -m pytest tests/test_head_gpu.py -x -q -k "bf16 or sharded or edge" 2>&1 | tail -5 >> $OUT
cat $OUT
