# GPU-box check: parity tests, smoke. Usage: gpurun -- 'bash scripts/gpu_check.sh <tag>'
TAG=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -15 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for p in inplace batched persistent; do
  QMC_ENERGY_PATH=$p python bench.py --config C2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c2_energy_${p}_$TAG.log 2>&1
  python - gpurun_out/c2_energy_${p}_$TAG.log $p <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln); print("C2 energy path", sys.argv[2], "%.1f k energies/s, sweep %.1f M/s" % (d["local_energies_per_s"] / 1e3, d["sweep_proposals_per_s"] / 1e6))
PY
done
