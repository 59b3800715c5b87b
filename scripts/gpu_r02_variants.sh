# round-2 sweep-kernel variants: parity under each variant, then a short C3 bench per variant.
# Usage: gpurun -- 'bash scripts/gpu_r02_variants.sh <tag>'
TAG=${1:-r02a}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --sweep-its 2000 --no-cpu-baseline"
pick() { python - "$1" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("sweep %.3f M/s  whole %.3f M/s  energies %.1f k/s  grad %.1f ms" % (
            d["sweep_proposals_per_s"] / 1e6, d["value"] / 1e6, d["local_energies_per_s"] / 1e3,
            d["segments_ms_per_step"]["gradient"]))
PY
}
QMC_IP_HS=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/pytest_hs_$TAG.log 2>&1; tail -3 gpurun_out/pytest_hs_$TAG.log
QMC_IP_HS=1 QMC_IP_LOCK=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_hslock_$TAG.log 2>&1; tail -3 gpurun_out/pytest_hslock_$TAG.log
for v in "base:" "hs:QMC_IP_HS=1" "lock1:QMC_IP_LOCK=1" "lock2:QMC_IP_LOCK=2" "hslock1:QMC_IP_HS=1 QMC_IP_LOCK=1" "hslock2:QMC_IP_HS=1 QMC_IP_LOCK=2" "hs_free:QMC_IP_HS=1 QMC_IP_SYNC=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs $B > gpurun_out/var_${name}_$TAG.log 2>&1
  echo "$name: $(pick gpurun_out/var_${name}_$TAG.log)"
done
