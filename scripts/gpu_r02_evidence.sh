# round-2 evidence run: default bench (both arms), ncu launch list of the same command, full ncu captures of the
# dominant kernel of every stage.  Usage: gpurun -- 'bash scripts/gpu_r02_evidence.sh <tag>'
TAG=${1:-r02}
mkdir -p gpurun_out
( time python bench.py --impl reference ) > gpurun_out/bench_ref_$TAG.log 2>&1
( time python bench.py ) > gpurun_out/bench_$TAG.log 2>&1
tail -4 gpurun_out/bench_$TAG.log | cut -c1-300
# launch list of the default step (one warm-up, one step): every kernel with its duration
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_list_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# full captures (one launch each) on the short profiling command: one full wave of 148 x 12 chains
CMD="python bench.py --steps 1 --warmup 1 --sweep-its 100 --chains 1776 --no-cpu-baseline"
$CMD > gpurun_out/plain_prof_$TAG.log 2>&1 || exit 1
for k in ${KERNELS:-k_sweep_ip:1 k_energy_ip:0 k_forward_plane:1 k_bwd_layerILi16ELi16:1 k_bwd_head:0}; do
  name=${k%%:*}; skip=${k#*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/${name}_$TAG -f $CMD > gpurun_out/ncu_full_${name}_$TAG.log 2>&1
done
ls -la gpurun_out/*_$TAG.ncu-rep
