# round-2 evidence run on one B200: tests, the default bench line, the reference arm, ncu launch list,
# single-pass DRAM traffic (warm L2 state, --cache-control none) and one --set full capture
set -u
mkdir -p gpurun_out
OUT=gpurun_out/final_n1.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv >> $OUT 2>&1
echo "== pytest -m gpu" >> $OUT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> $OUT
echo "== bench (default flags)" >> $OUT
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>> $OUT
tail -c 600 gpurun_out/r2_bench_n1.err >> $OUT
echo "== reference arm" >> $OUT
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> $OUT
echo "== ncu launch list" >> $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-cfg4 > gpurun_out/r2_launches.log 2>&1
echo "== ncu DRAM traffic, single pass, L2 state carried over" >> $OUT
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 44 -c 21 --csv \
  --log-file gpurun_out/r2_traffic.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-cfg4 > gpurun_out/r2_traffic.log 2>&1
echo "== ncu --set full" >> $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:umma_kernel|prep_kernel" -s 20 -c 5 \
  -f -o gpurun_out/r2_full_final python bench.py --no-cpu-baseline --no-graph --no-cfg4 --steps 2 --warmup 3 > gpurun_out/r2_full_final.log 2>&1
ls -la gpurun_out/r2_* >> $OUT
python - <<'PY' >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d["value_path"], {k:round(v*1000,1) for k,v in d["ms_per_step_by_path"].items()})
print("e2e", round(d["e2e"]["value"]), d["e2e"]["path"], {k:round(v*1000,1) for k,v in d["e2e"]["ms_per_step_by_path"].items()})
print("kernels", {k["kernel"][:10]:(round(k["ms"]*1000,1), round(k.get("frac",0),3)) for k in d["kernels"]})
print("roofline", d["roofline"])
print("cfg4", d.get("cfg4",{}).get("value"), d.get("cfg4",{}).get("ms_per_step"))
print("clocks", d["clocks"], "cpu", d.get("cpu_baseline"))
PY
cat $OUT
