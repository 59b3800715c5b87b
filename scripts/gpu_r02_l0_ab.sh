# k_sweep_ip, layer-0 site decode: FastDiv with constant-memory magics (in-tree) against integer division (the code of the
# r02d evidence run; variants/libqmc_l0div.so = -DQMC_IP_L0_FASTDIV=0), alternating, full-length C3 steps.
# Usage: gpurun -- 'bash scripts/gpu_r02_l0_ab.sh <tag>'
TAG=${1:-r02i}
mkdir -p gpurun_out
out=gpurun_out/l0_ab_$TAG.txt; : > $out
one() { python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" 2>&1 | grep '^{' | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
print('sweep %.4f M/s  value %.4f M/s  frac %.4f' % (d['sweep_proposals_per_s'] / 1e6, d['value'] / 1e6, d['roofline']['frac']))"; }
for i in 1 2; do
  echo "fastdiv  $(one)" >> $out
  echo "division $(one --lib variants/libqmc_l0div.so)" >> $out
done
cat $out
