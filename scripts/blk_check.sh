# column-blocked Wb + dW clip fix: parity tests, then the default bench
set -u
mkdir -p gpurun_out
OUT=gpurun_out/blk_check.txt
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -25 >> $OUT
echo "== bench" >> $OUT
timeout 900 python bench.py > gpurun_out/blk_bench_n1.json 2> gpurun_out/blk_bench_n1.err
tail -c 400 gpurun_out/blk_bench_n1.err >> $OUT
python - <<'PY' >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/blk_bench_n1.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d["value_path"], {k:round(v*1000,1) for k,v in d["ms_per_step_by_path"].items()})
print("e2e", round(d["e2e"]["value"]), d["e2e"]["path"])
print("kernels", {k["kernel"][:10]:(round(k["ms"]*1000,1), round(k.get("frac",0),3)) for k in d["kernels"]})
print("phases", [(p["phase"], round(p["ms"]*1000,1)) for p in d.get("phases",[])])
print("cfg4", d.get("cfg4",{}).get("ms_per_step"))
PY
