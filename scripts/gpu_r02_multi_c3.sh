# 8-GPU line of the default workload (C3, 32768 chains).  Usage: gpurun --gpus 8 -- 'bash scripts/gpu_r02_multi_c3.sh <tag>'
TAG=${1:-r02}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/multi_C3_$TAG.log 2>&1
grep '^{' gpurun_out/multi_C3_$TAG.log | tail -1 > gpurun_out/bench_line_multi_C3_$TAG.json
cut -c1-250 gpurun_out/bench_line_multi_C3_$TAG.json
