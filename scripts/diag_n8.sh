set -u
N=${1:-8}
mkdir -p gpurun_out
OUT=gpurun_out/diag_scale$N.txt
: > $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_check.py nvlink 2>&1 | grep -E "rank 0|FAIL|Error|error" | tail -8 >> $OUT
run() {  # name workload env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 50 --warmup 5 --workload $wl 2>gpurun_out/scale_$name.err | tail -1 > gpurun_out/scale_$name.json
  python - <<PY >> $OUT 2>&1
import json
try:
    d=json.loads(open("gpurun_out/scale_$name.json").read())
    print("$name N=$N", round(d["value"]), round(d["ms_per_step"]*1000,1), "us", d.get("value_path"), "e2e", round(d["e2e"]["value"]), {k:round(v*1000,1) for k,v in d["ms_per_step_by_path"].items()}, {k["kernel"][:10]:round(k["ms"]*1000,1) for k in d["kernels"]})
except Exception as e:
    print("$name FAILED", e)
PY
}
run cfg3_bn128 cfg3 ASM_UMMA_BN=128
run cfg3_bn128_pdlgraph cfg3 ASM_UMMA_BN=128 ASM_PDL_GRAPH=1
run cfg3_bn256_pdlgraph cfg3 ASM_PDL_GRAPH=1
cat $OUT
