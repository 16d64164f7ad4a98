# multi-GPU check: parity of both transports on N ranks, then the N-rank bench line
set -u
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/diag_n$N.txt
: > $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_check.py 2>&1 | grep -E "rank 0|FAIL|Error|error" | tail -20 >> $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 50 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read())
print("N=$N", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d.get("value_path"), "e2e", round(d["e2e"]["value"]), {k:round(v*1000,1) for k,v in d["ms_per_step_by_path"].items()}, {k["kernel"][:10]:round(k["ms"]*1000,1) for k in d["kernels"]})
PY
tail -5 gpurun_out/bench_n$N.err >> $OUT
cat $OUT
