# 8-GPU lines: C3 (32768 chains) and C5 (8192 chains per GPU, full VMC gradient step).  Usage: gpurun --gpus 8 -- 'bash scripts/gpu_r02_multi.sh <tag>'
TAG=${1:-r02}
mkdir -p gpurun_out
: > gpurun_out/bench_lines_multi_gpu_$TAG.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/multi_C3_$TAG.log 2>&1
grep '^{' gpurun_out/multi_C3_$TAG.log | tail -1 >> gpurun_out/bench_lines_multi_gpu_$TAG.json
$TR bench.py --gpus 8 --config C5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/multi_C5_$TAG.log 2>&1
grep '^{' gpurun_out/multi_C5_$TAG.log | tail -1 >> gpurun_out/bench_lines_multi_gpu_$TAG.json
cut -c1-250 gpurun_out/bench_lines_multi_gpu_$TAG.json
