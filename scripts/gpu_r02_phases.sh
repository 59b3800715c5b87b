# phase-cycle breakdown of k_sweep_ip (variants/lib_prof.so = scripts/build_variant.sh ... "-DQMC_IP_PROFILE=1") at
# 1, 2 and 3 warps per scheduler.  Usage: gpurun -- 'bash scripts/gpu_r02_phases.sh <tag> [lib]'
TAG=${1:-r02p}
LIB=${2:-variants/lib_prof.so}
mkdir -p gpurun_out
for w in 4 8 12; do
  echo "== $w warps per SM" >> gpurun_out/phases_$TAG.log
  python bench.py --steps 2 --warmup 1 --sweep-its 600 --chains $((148 * w)) --no-cpu-baseline --lib $LIB \
      --tuning flags=4,max_warps=$w 2>> gpurun_out/phases_$TAG.log | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('sweep %.3f M/s' % (d['sweep_proposals_per_s']/1e6))" >> gpurun_out/phases_$TAG.log
done
cat gpurun_out/phases_$TAG.log
