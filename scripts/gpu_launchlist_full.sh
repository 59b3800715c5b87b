# ncu launch list of the DEFAULT bench command (the one the driver runs), for profiles/.
set -x
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/bench_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_full.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_list_full.log 2>&1
tail -1 gpurun_out/bench_plain_full.log | cut -c1-300
