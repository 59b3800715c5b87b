# time-slicing granularity of k_sweep_ip (max chunks per chain), C3 default bench
mkdir -p gpurun_out
for c in 16 32 128; do
  QMC_IP_CHUNKS=$c timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/chunks_$c.log 2>&1
  echo "chunks $c: $(tail -1 gpurun_out/chunks_$c.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["gpu_launches"])' 2>&1 | tail -1)"
done
