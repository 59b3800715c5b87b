# full-length bench (both arms) + ncu evidence for profiles/. Usage: gpurun -- 'bash scripts/gpu_bench_full.sh <tag>'
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.log 2>&1
( time python bench.py ) > gpurun_out/bench_$TAG.log 2>&1
tail -4 gpurun_out/bench_$TAG.log | cut -c1-400
CMD="python bench.py --steps 1 --warmup 1 --sweep-its 100 --chains 1036 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/sweep_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
