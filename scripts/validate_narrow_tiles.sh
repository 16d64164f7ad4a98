# Validation + A/B of the opt-in 128-wide tiles (ASM_UMMA_BN=128).
# Run on a B200 box:   bash scripts/validate_narrow_tiles.sh          (1 GPU)
#                      bash scripts/validate_narrow_tiles.sh 8        (8 GPUs, scaling A/B)
# Step 1 runs the whole GPU parity suite with the narrow tiles forced on; step 2 times both.
set -u
N=${1:-1}
mkdir -p gpurun_out
echo "== parity suite with ASM_UMMA_BN=128"
ASM_UMMA_BN=128 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for bn in 256 128 256 128; do
  if [ "$N" = "1" ]; then
    ASM_UMMA_BN=$bn timeout 200 python bench.py --no-cpu-baseline --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bn_$bn.json
  else
    ASM_UMMA_BN=$bn timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bn_$bn.json
  fi
  python - <<PY
import json
d = json.loads(open("gpurun_out/bn_$bn.json").read())
print("BN=$bn N=$N", round(d["value"]), "samples/s", round(d["ms_per_step"] * 1000, 1), "us", d.get("value_path"),
      {k["kernel"][:6]: round(k["ms"] * 1000, 1) for k in d.get("kernels", [])})
PY
done

# Other opt-ins waiting for a measurement (same A/B harness):
#   bash scripts/ab_bench.sh BENCH_LOSS_STREAM "0 1"        # e2e: loss read-back on its own stream
#   ASM_PREP_AUTO=1 at N = 4 / 8                            # norm-kernel shape for small shards
#   ASM_OPT_STREAM=1 python scripts/bench_train_step.py     # streaming optimizer vs fused epilogue
#   ASM_OPT_STREAM=1 python -m pytest tests -q -m gpu -k optimizer
