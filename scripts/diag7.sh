set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag7.txt
: > $OUT
echo "== gpu tests" >> $OUT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 >> $OUT
echo "== optimizer tests with the streaming update (ASM_OPT_STREAM=1)" >> $OUT
ASM_OPT_STREAM=1 timeout 600 python -m pytest tests/test_head_gpu.py -x -q -m gpu -k "optimizer" 2>&1 | tail -3 >> $OUT
echo "== bench default (cfg3 + cfg4 sub-record)" >> $OUT
( time BENCH_VERBOSE=1 timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/diag7_bench.json 2> gpurun_out/diag7_bench.err ) 2>> $OUT
grep "bench " gpurun_out/diag7_bench.err | tail -12 >> $OUT
python - <<'PY' >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/diag7_bench.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d["value_path"], "e2e", round(d["e2e"]["value"]), d["e2e"]["path"], {k:round(v*1000,1) for k,v in d["e2e"]["ms_per_step_by_path"].items()})
print("parity", d["parity"]["ok"], d["parity"]["paths"])
print("kernels", {k["kernel"][:10]:round(k["ms"]*1000,1) for k in d["kernels"]}, "roofline", d["roofline"])
print("cfg4", d.get("cfg4", {}).get("value"), d.get("cfg4", {}).get("ms_per_step"), d.get("cfg4", {}).get("parity", {}).get("paths"))
print("cpu", d.get("cpu_baseline"))
PY
echo "== bench cfg5" >> $OUT
timeout 600 python bench.py --steps 20 --warmup 3 --workload cfg5 --no-cpu-baseline > gpurun_out/diag7_cfg5.json 2> gpurun_out/diag7_cfg5.err
tail -3 gpurun_out/diag7_cfg5.err >> $OUT
python - <<'PY' >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/diag7_cfg5.json").read().strip().splitlines()[-1])
print("cfg5", d.get("value"), d.get("ms_per_step"), d.get("parity"))
PY
cat $OUT
