# GPU-box check: parity tests, smoke, short bench lines. Usage: gpurun -- 'bash scripts/gpu_check.sh <tag>'
TAG=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -15 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in C3 C1 C2 C4; do
  python bench.py --config $c --steps 2 --warmup 1 --sweep-its 2000 --no-cpu-baseline > gpurun_out/short_${c}_$TAG.log 2>&1
  tail -1 gpurun_out/short_${c}_$TAG.log | cut -c1-330
done
