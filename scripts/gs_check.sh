set -u
mkdir -p gpurun_out
OUT=gpurun_out/gs_check.txt
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -25 >> $OUT
for v in "ASM_BWD_STREAM=1" "ASM_BWD_STREAM=0" "ASM_BWD_STREAM=1"; do
  echo "== bench $v" >> $OUT
  env $v timeout 600 python bench.py --no-cpu-baseline --no-cfg4 2> gpurun_out/gs_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg3', round(d['ms_per_step']*1000,1), d['value_path'], {k:round(v*1000,1) for k,v in d['ms_per_step_by_path'].items()}, 'parity', d['parity']['ok'], {k['kernel'][:10]:round(k['ms']*1000,1) for k in d['kernels']}, [(p['phase'][:8], round(p['ms']*1000,1)) for p in d.get('phases',[])])
print('  parity', d['parity']['paths'])" >> $OUT 2>&1
  tail -c 300 gpurun_out/gs_bench.err >> $OUT
done
