mkdir -p gpurun_out
for dbg in 0 8 4 0 8 4; do
  ASM_UMMA_DEBUG=$dbg timeout 100 python bench.py --no-cpu-baseline --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/ab_$dbg.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$dbg.json").read())
print("dbg=$dbg", round(d["value"]), round(d["ms_per_step"]*1000,1), d["clocks"]["sm_mhz"], {k["kernel"][:6]:round(k["ms"]*1000,1) for k in d["kernels"]})
PY
done
