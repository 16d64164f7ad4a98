mkdir -p gpurun_out
( timeout 70 python bench.py --workload cfg5 --no-cpu-baseline --no-graph --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg5', round(d['value']), round(d['ms_per_step']*1000,1), d['parity']['ok'], d['parity']['paths'])"
  timeout 70 python bench.py --mode fp32 --no-cpu-baseline --no-graph --no-cfg4 --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fp32', round(d['value']), round(d['ms_per_step']*1000,1), d['parity']['ok'], d['parity']['paths'])" ) > gpurun_out/last_check.txt 2>&1
