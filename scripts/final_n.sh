# round-2 evidence run on N B200s: multi-GPU parity of both transports, then the bench line (with the cfg4 sub-record)
set -u
N=${1:-8}
mkdir -p gpurun_out
OUT=gpurun_out/final_n$N.txt
: > $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_check.py 2>&1 | grep -E "rank 0|FAIL|Error|error" | tail -14 >> $OUT
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2>> $OUT
grep -E "NCCL INFO (comm|Init|Connected)|nranks|Error|error|Traceback" gpurun_out/r2_bench_n$N.err | head -6 >> $OUT
python - <<PY >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N cfg3", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d["value_path"], {k:round(v*1000,1) for k,v in d["ms_per_step_by_path"].items()})
print("e2e", round(d["e2e"]["value"]), d["e2e"]["path"], {k:round(v*1000,1) for k,v in d["e2e"]["ms_per_step_by_path"].items()})
print("parity", d["parity"]["ok"], {k:(round(v["loss_rel"],8), round(v["cos_dX"],6)) for k,v in d["parity"]["paths"].items()})
print("kernels", {k["kernel"][:12]:round(k["ms"]*1000,1) for k in d["kernels"]}, "launches/step", d.get("launches_per_step"), d.get("transport"))
c4=d.get("cfg4",{})
print("cfg4", c4.get("value"), c4.get("ms_per_step"), c4.get("value_path"), c4.get("parity",{}).get("ok"))
PY
cat $OUT
