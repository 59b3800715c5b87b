# The parity-test configurations (C1, C2, C4) with the final build: JSON lines at N = 1 (full length) and one full ncu
# capture of the sweep kernel each of C2 (classic persistent kernel) and C4 (k_sweep_sym) runs, so that their roofline
# rows in profiles/r02_summary.md rest on measurements.  Usage: gpurun -- 'bash scripts/gpu_r02_small_configs.sh <tag>'
TAG=${1:-r02e}
mkdir -p gpurun_out
: > gpurun_out/bench_lines_small_$TAG.json
run() { name=$1; shift; python bench.py "$@" --no-cpu-baseline > gpurun_out/cfg_${name}_$TAG.log 2>&1; grep '^{' gpurun_out/cfg_${name}_$TAG.log | tail -1 >> gpurun_out/bench_lines_small_$TAG.json; tail -1 gpurun_out/cfg_${name}_$TAG.log | cut -c1-240; }
run C1 --config C1 --steps 3 --warmup 3
run C2 --config C2 --steps 3 --warmup 3
run C4 --config C4 --steps 2 --warmup 1
# launch lists (which kernels, how long) and one full capture of the sweep kernel of each
for c in ${CAPTURES:-C2 C4}; do
  CMD="python bench.py --config $c --steps 1 --warmup 1 --sweep-its 200 --no-cpu-baseline"
  $CMD > gpurun_out/plain_${c}_$TAG.log 2>&1 || { echo "plain run of $c failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${c}_$TAG.csv $CMD > gpurun_out/ncu_list_${c}_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k 'regex:k_sweep' -s 1 -c 1 -o gpurun_out/k_sweep_${c}_$TAG -f $CMD > gpurun_out/ncu_full_${c}_$TAG.log 2>&1
done
ls -la gpurun_out/*_$TAG.ncu-rep
