set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag9.txt
: > $OUT
echo "== gpu tests" >> $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $OUT
for i in 1 2; do timeout 200 python scripts/kernel_times.py 512 512 85742 40 2>&1 | tail -1 >> $OUT; done
timeout 200 python scripts/kernel_times.py 2048 512 85742 20 2>&1 | tail -1 >> $OUT
timeout 120 python scripts/center_times.py 2>&1 | tail -2 >> $OUT
echo "== bench cfg5" >> $OUT
timeout 600 python bench.py --steps 20 --warmup 3 --workload cfg5 --no-cpu-baseline 2>gpurun_out/diag9_cfg5.err | tail -1 > gpurun_out/diag9_cfg5.json
python - <<'PY' >> $OUT 2>&1
import json
d=json.loads(open("gpurun_out/diag9_cfg5.json").read())
print("cfg5", round(d["value"]), round(d["ms_per_step"]*1000,1), "us parity", d["parity"]["ok"], d["parity"]["paths"])
PY
echo "== cfg2 end to end: SphereFaceNet-20 + head, batch 512, C=10572 (train.py:231-239 images/s)" >> $OUT
timeout 600 python examples/train_sphereface20.py --steps 40 --batch 512 --classes 10572 2>&1 | tail -6 >> $OUT
cat $OUT
