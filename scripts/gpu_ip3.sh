mkdir -p gpurun_out
for sy in 1 2; do
QMC_IP_SYNC=$sy timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "inplace" 2>&1 | tail -2
done
for sy in 0 1 2; do
  for w in 12 9; do
  QMC_IP_SYNC=$sy QMC_MAX_WARPS=$w QMC_SWEEP_PATH=inplace timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 1000 --no-cpu-baseline > gpurun_out/ip_s${sy}_w$w.log 2>&1
  echo "inplace sync$sy w$w: $(tail -1 gpurun_out/ip_s${sy}_w$w.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"])' 2>&1 | tail -1)"
  done
done
