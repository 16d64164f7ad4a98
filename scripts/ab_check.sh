set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab_check.txt
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -15 >> $OUT
for v in "A=1" "ASM_B200_LIB=tf_face_toolbox_b200/lib/alt/libasoftmax_b200.so" "A=1" "ASM_B200_LIB=tf_face_toolbox_b200/lib/alt/libasoftmax_b200.so"; do
  echo "== bench $v" >> $OUT
  env $v timeout 600 python bench.py --no-cpu-baseline --no-cfg4 --no-graph 2> gpurun_out/ab_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg3', round(d['ms_per_step']*1000,1), 'parity', d['parity']['ok'], {k['kernel'][:10]:round(k['ms']*1000,1) for k in d['kernels']}, [(p['phase'][:8], round(p['ms']*1000,1)) for p in d.get('phases',[])])" >> $OUT 2>&1
  tail -c 200 gpurun_out/ab_bench.err >> $OUT
done
