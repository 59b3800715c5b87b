# Round-end evidence run (one GPU):  gpurun --timeout 1500 -- 'bash scripts/gpu_profile_final.sh r01ip'
#  1. the default bench (both arms) without a profiler  -> gpurun_out/bench_<tag>.log, bench_ref_<tag>.log
#  2. ncu launch list of the SAME default command       -> gpurun_out/launches_<tag>.csv
#  3. one --set full capture of the dominant kernel     -> gpurun_out/sweep_<tag>.ncu-rep
TAG=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err && tail -1 gpurun_out/bench_$TAG.log | cut -c1-400
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-300
python bench.py --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --sweep-its 100 --chains 1776 --no-cpu-baseline"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/sweep_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
