# proposals/s of the two sweep decompositions against chains per GPU (C3, short sweeps)
mkdir -p gpurun_out
for ch in 4096 8192 16384 32768; do
  for path in persistent batched; do
    QMC_SWEEP_PATH=$path timeout 300 python bench.py --config C3 --chains $ch --steps 2 --warmup 3 --sweep-its 1000 --no-cpu-baseline > gpurun_out/chains_${path}_${ch}.log 2>&1
    echo "$path $ch: $(tail -1 gpurun_out/chains_${path}_${ch}.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["sweep_proposals_per_s"], d["local_energies_per_s"], d["segments_ms_per_step"])' 2>&1 | tail -1)"
  done
done
