# parameter-block staging: TMA bulk copy + mbarrier (default) against the per-thread copy loop (-DQMC_PARAMS_TMA=0)
mkdir -p gpurun_out
for lib in qmcnn_b200/libqmcnn_b200.so variants/lib_notma.so; do
  echo "== $lib"
  python scripts/time_kernels.py C3 4096 0 $PWD/$lib
  python scripts/time_kernels.py C3 256 0 $PWD/$lib
  python bench.py --steps 2 --warmup 1 --sweep-its 2000 --no-cpu-baseline --lib $lib | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('C3 short step: sweep %.3f M/s, energies %.1f k/s, gradient %.2f ms' % (d['sweep_proposals_per_s']/1e6, d['local_energies_per_s']/1e3, d['segments_ms_per_step']['gradient']))"
done
