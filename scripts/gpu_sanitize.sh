# compute-sanitizer on the smallest cases (one tool per gpurun call). Usage: gpurun -- 'bash scripts/gpu_sanitize.sh memcheck|racecheck|initcheck|synccheck'
set -x
TOOL=${1:-memcheck}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import numpy as np, torch
import __graft_entry__ as g
g.smoke()                                   # CRBM: forward, classic sweep, energy (persistent), backward
import qmcnn_b200 as q
torch.manual_seed(0)
m = q.DCRBM(3, [16, 16, 8], 2, seed=1); m.flat.mul_(10)
S = type("S", (q.Sampler,), dict(MAX_NUM_SAMPLERS=10**9, SWEEPFACTOR=1, THERMFACTOR=1))
for env in ({}, {"QMC_SWEEP_PATH": "batched"}, {"QMC_LEAN": "1"}):
    os.environ.update(env)
    mm = q.DCRBM(3, [16, 16, 8], 2, seed=1); mm.flat.copy_(m.flat)
    s = S(mm, (8, 9), 7, 12, 1, seed=3)
    out = s.mcmc_op()
    e = q.ising_energy(mm, out, system_shape=(8, 9), H=1.0)      # batched energy path
    eh = q.heisenberg_energy(mm, out, system_shape=(8, 9))       # persistent pair-box path
    gr = q.logpsi_gradient(mm, out, (e - e.mean()) / e.numel(), (8, 9))
    print(env, float(e.real.mean()), float(eh.real.mean()), float(gr.abs().max()), s.acceptance_count)
    for k in env: os.environ.pop(k)
sm = q.SymmetrizedModel(q.CRBM(3, 1, 2, 2, seed=2)); sm.base.flat.mul_(20)
ss = S(sm, (6, 6), 3, 8, 2, seed=5); o = ss.mcmc_op()
print("sym", float(q.heisenberg_energy(sm, o, system_shape=(6, 6)).real.mean()), ss.acceptance_count)
torch.cuda.synchronize()
PY
compute-sanitizer --tool $TOOL --error-exitcode 9 python /tmp/san_case.py > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "exit=$?" >> gpurun_out/sanitizer_$TOOL.log
tail -12 gpurun_out/sanitizer_$TOOL.log
