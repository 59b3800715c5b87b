# one full ncu capture of the in-place sweep kernel (k_sweep_ip).
# Usage: gpurun -- 'bash scripts/gpu_profile_ip.sh <tag> [chains] [extra bench.py arguments]'
TAG=${1:-ip}
CH=${2:-1776}
shift; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --sweep-its 100 --chains $CH --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/sweep_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
