# phase-group start offsets for k_sweep_ip: throughput + conv concurrency.  Usage: gpurun -- 'bash scripts/gpu_r02_stagger.sh <tag> <lib> "<kcycles list>"'
TAG=${1:-r02s}
LIB=${2:-variants/lib_prof.so}
mkdir -p gpurun_out
for st in ${3:-0 29 57 86}; do
  echo "== stagger $st x 1024 cycles" >> gpurun_out/stagger_$TAG.log
  python bench.py --steps 2 --warmup 1 --sweep-its 2000 --no-cpu-baseline --lib $LIB \
      --tuning ip_stagger=$st 2>> gpurun_out/stagger_$TAG.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('sweep %.3f M/s' % (d['sweep_proposals_per_s']/1e6))" >> gpurun_out/stagger_$TAG.log
done
cat gpurun_out/stagger_$TAG.log
