mkdir -p gpurun_out
for g in 2 3 6; do
  QMC_IP_SYNC=3 QMC_IP_GROUP=$g QMC_SWEEP_PATH=inplace timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 2000 --no-cpu-baseline > gpurun_out/ip6_g${g}.log 2>&1
  echo "inplace sync3 group$g: $(tail -1 gpurun_out/ip6_g${g}.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["gpu_launches"])' 2>&1 | tail -1)"
done
for w in 11 10; do
  QMC_IP_SYNC=3 QMC_MAX_WARPS=$w QMC_SWEEP_PATH=inplace timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 2000 --no-cpu-baseline > gpurun_out/ip6_w${w}.log 2>&1
  echo "inplace sync3 warps$w: $(tail -1 gpurun_out/ip6_w${w}.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["gpu_launches"])' 2>&1 | tail -1)"
done
