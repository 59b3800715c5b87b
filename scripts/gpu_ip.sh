# in-place sweep kernel: parity against the classic kernel, then sweep throughput of both
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "inplace or lean or batched_path" 2>&1 | tail -5
for path in pingpong inplace; do
  QMC_SWEEP_PATH=$path timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 1000 --no-cpu-baseline > gpurun_out/ip_${path}.log 2>&1
  echo "$path: $(tail -1 gpurun_out/ip_${path}.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["segments_ms_per_step"], d["roofline"]["frac"])' 2>&1 | tail -1)"
done
for w in 8 10 11; do
  QMC_MAX_WARPS=$w QMC_SWEEP_PATH=inplace timeout 300 python bench.py --config C3 --steps 2 --warmup 3 --sweep-its 1000 --no-cpu-baseline > gpurun_out/ip_w$w.log 2>&1
  echo "inplace w$w: $(tail -1 gpurun_out/ip_w$w.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"])' 2>&1 | tail -1)"
done
