# GPU-box check: parity tests, smoke, a short bench. Usage: gpurun -- 'bash scripts/gpu_check.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
