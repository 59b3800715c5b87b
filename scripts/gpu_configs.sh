# short benches of the other BASELINE configs (parity-test cases, not the headline)
set -x
mkdir -p gpurun_out
python bench.py --config C1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C1.log 2>&1
python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.log 2>&1
python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4.log 2>&1
python bench.py --config C5 --steps 2 --warmup 3 --sweep-its 2000 --no-cpu-baseline > gpurun_out/bench_C5.log 2>&1
QMC_SWEEP_PATH=batched python bench.py --config C5 --steps 2 --warmup 3 --sweep-its 2000 --no-cpu-baseline > gpurun_out/bench_C5_batched.log 2>&1
for c in C1 C2 C4 C5 C5_batched; do tail -1 gpurun_out/bench_$c.log | cut -c1-200; done
