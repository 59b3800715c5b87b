# every BASELINE config as a recorded JSON line at N = 1 (full length), plus C3 at sigma = 0.1 (a non-trivial acceptance
# rate).  Usage: gpurun -- 'bash scripts/gpu_r02_configs.sh <tag>'
TAG=${1:-r02}
mkdir -p gpurun_out
: > gpurun_out/bench_lines_configs_$TAG.json
run() { name=$1; shift; python bench.py "$@" --no-cpu-baseline > gpurun_out/cfg_${name}_$TAG.log 2>&1; grep '^{' gpurun_out/cfg_${name}_$TAG.log | tail -1 >> gpurun_out/bench_lines_configs_$TAG.json; tail -1 gpurun_out/cfg_${name}_$TAG.log | cut -c1-200; }
run C1 --config C1 --steps 3 --warmup 3
run C2 --config C2 --steps 3 --warmup 3
run C3_sigma0.1 --config C3 --steps 2 --warmup 1 --sigma 0.1
run C4 --config C4 --steps 2 --warmup 1
run C5 --config C5 --steps 2 --warmup 1
# the weight-gradient / cotangent kernel of the hidden layers (full capture, profiling command)
CMD="python bench.py --steps 1 --warmup 1 --sweep-its 100 --chains 1776 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k 'regex:^k_bwd_layer$' -s 1 -c 1 -o gpurun_out/k_bwd_layer_$TAG -f $CMD > gpurun_out/ncu_full_k_bwd_layer_$TAG.log 2>&1
ls -la gpurun_out/k_bwd_layer_$TAG.ncu-rep
