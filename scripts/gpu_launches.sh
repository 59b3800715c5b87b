# per-kernel durations of a short bench. Usage: gpurun -- 'bash scripts/gpu_launches.sh <tag> [bench args]'
set -x
TAG=${1:-x}; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
