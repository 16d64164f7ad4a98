set -u
mkdir -p gpurun_out
OUT=gpurun_out/diag6.txt
: > $OUT
timeout 60 python scripts/stream_probe.py >> $OUT 2>&1
env | grep -i -E "cuda|nvidia" >> $OUT
nvidia-smi -q | grep -i -E "compute mode|mig mode|persistence" >> $OUT
cat $OUT
