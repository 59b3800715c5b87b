# round-2 sweep-kernel variants: parity, then a short C3 bench per variant (env knobs are experiment-only).
# Usage: gpurun -- 'bash scripts/gpu_r02_variants.sh <tag> "name:ENV=..;name2:ENV=.."'
TAG=${1:-r02a}
VARS=${2:-"base:"}
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
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
IFS=';' read -ra VV <<< "$VARS"
for v in "${VV[@]}"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs $B > gpurun_out/var_${name}_$TAG.log 2>&1
  echo "$name: $(pick gpurun_out/var_${name}_$TAG.log)"
done
