# streaming prep side by side with the forward GEMM: parity tests, then A/B bench
set -u
mkdir -p gpurun_out
OUT=gpurun_out/stream_check.txt
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -25 >> $OUT
for v in 0 1; do
  echo "== bench ASM_PREP_STREAM=$v" >> $OUT
  ASM_PREP_STREAM=$v timeout 600 python bench.py --no-cpu-baseline --no-cfg4 > gpurun_out/stream_bench_$v.json 2> gpurun_out/stream_bench_$v.err
  tail -c 300 gpurun_out/stream_bench_$v.err >> $OUT
  python - $v <<'PY' >> $OUT 2>&1
import json,sys
d=json.loads(open("gpurun_out/stream_bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print("cfg3", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d["value_path"], {k:round(v*1000,1) for k,v in d["ms_per_step_by_path"].items()}, "parity", d["parity"]["ok"])
print("e2e", round(d["e2e"]["value"]), d["e2e"]["path"])
PY
done
echo "== bench ASM_PREP_STREAM=0 ASM_UMMA_BN=128" >> $OUT
ASM_PREP_STREAM=0 ASM_UMMA_BN=128 timeout 600 python bench.py --no-cpu-baseline --no-cfg4 --no-graph 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg3 bn128', round(d['ms_per_step']*1000,1), {k['kernel'][:10]:round(k['ms']*1000,1) for k in d['kernels']})" >> $OUT 2>&1
