# Classic persistent kernels (C1 / C2 / C4: k_sweep_w8 / w16, k_sweep_sym, k_energy_w*): the in-tree build (constant-memory
# division magics, tanh fast path, small task counts spread over the SMs) and the variant builds with more warps per CTA for
# small models (scripts/build_variant.sh variants/libqmc_w28.so "-DQMC_SMALL_WARPS=28", ... 24, 20), against the lines of
# scripts/gpu_r02_small_configs.sh (the build before).  Also: golden chains bit-exact with the in-tree build, and the C3
# line (k_sweep_ip shares qmc_device.cuh).  Usage: gpurun -- 'bash scripts/gpu_r02_classic_ab.sh <tag>'
TAG=${1:-r02f}
mkdir -p gpurun_out
out=gpurun_out/classic_ab_$TAG.txt
lines=gpurun_out/classic_ab_lines_$TAG.json
: > $out; : > $lines
summ() { python -c "
import json, sys
l = sys.stdin.readline()
if not l.startswith('{'): print('  (no line)'); sys.exit(0)
d = json.loads(l)
print('  %s value %.4g M/s  sweep %.4g M/s  energies %.4g /s  ms %s' % (d['config']['workload'][:3], d['value'] / 1e6,
      d.get('sweep_proposals_per_s', 0) / 1e6, d.get('local_energies_per_s', 0), d.get('segments_ms_per_step')))"; }
line() { python bench.py "$@" --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee -a $lines | summ >> $out; }
python -m pytest tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/golden_$TAG.log 2>&1; tail -1 gpurun_out/golden_$TAG.log | tee -a $out
for v in tree w28 w24; do
  lib=""; [ $v != tree ] && lib="--lib variants/libqmc_$v.so"
  [ $v != tree ] && [ ! -f variants/libqmc_$v.so ] && continue
  echo "== $v" >> $out
  line --config C2 --steps 3 --warmup 3 $lib
  [ $v != tree ] && continue             # QMC_SMALL_WARPS only changes k_sweep_w16's launch
  line --config C1 --steps 3 --warmup 3
  line --config C4 --steps 1 --warmup 1
done
echo "== tree, C3 (k_sweep_ip)" >> $out
line --config C3 --steps 2 --warmup 1
cat $out
