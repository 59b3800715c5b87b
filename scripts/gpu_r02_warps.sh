# per-warp proposal latency of k_sweep_ip as a function of warps per SM (one full wave per launch: chains = 148 * warps).
# Usage: gpurun -- 'bash scripts/gpu_r02_warps.sh <tag>'
TAG=${1:-r02w}
mkdir -p gpurun_out
for w in 4 8 12; do
  for grp in 4; do
    python bench.py --steps 2 --warmup 1 --sweep-its 600 --chains $((148 * w)) --no-cpu-baseline \
        --tuning flags=4,max_warps=$w,ip_group=$grp > gpurun_out/warps_${w}_g${grp}_$TAG.log 2>&1
    python - gpurun_out/warps_${w}_g${grp}_$TAG.log $w $grp <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln); w = int(sys.argv[2])
        r = d["sweep_proposals_per_s"]
        print("warps %2d group %s: sweep %.3f M/s = %.1f k cycles per proposal per warp (1.965 GHz)" % (
            w, sys.argv[3], r / 1e6, 148 * w / r * 1.965e9 / 1e3))
PY
  done
done
