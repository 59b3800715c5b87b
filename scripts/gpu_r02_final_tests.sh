# Final state of the round on a B200: the whole GPU suite against the production build and against the debug build
# (device-side bounds checks), smoke(), and the C1 / C2 / C4 lines + one full capture of C2's sweep kernel.
# Usage: gpurun -- 'bash scripts/gpu_r02_final_tests.sh <tag>'
TAG=${1:-r02g}
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
( time python -m pytest tests -x -q -m gpu --qmc-lib qmcnn_b200/libqmcnn_b200_debug.so ) > gpurun_out/pytest_gpu_debug_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_debug_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
rm -rf variants
bash scripts/gpu_r02_small_configs.sh $TAG
