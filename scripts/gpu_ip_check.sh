mkdir -p gpurun_out
for sy in 0 3; do
QMC_IP_SYNC=$sy timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -k "inplace or golden" 2>&1 | tail -1
done
QMC_IP_SYNC=3 QMC_SWEEP_PATH=inplace timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 2000 --no-cpu-baseline > gpurun_out/ip9.log 2>&1
echo "inplace sync3: $(tail -1 gpurun_out/ip9.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["gpu_launches"])' 2>&1 | tail -1)"
QMC_IP_SYNC=3 bash scripts/gpu_profile_ip.sh ${1:-ip9}
