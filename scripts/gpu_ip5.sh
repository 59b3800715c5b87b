mkdir -p gpurun_out
for sy in 0 3; do
QMC_IP_SYNC=$sy timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -k "inplace or golden" 2>&1 | tail -2
done
for sy in 0 3 2; do
  QMC_IP_SYNC=$sy QMC_SWEEP_PATH=inplace timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 2000 --no-cpu-baseline > gpurun_out/ip5_s${sy}.log 2>&1
  echo "inplace sync$sy: $(tail -1 gpurun_out/ip5_s${sy}.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["gpu_launches"])' 2>&1 | tail -1)"
done
