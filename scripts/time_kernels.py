"""Time K1 (forward) and K4 (gradient) alone with CUDA events and print them against the FP32 roofline.
Usage: python scripts/time_kernels.py [C3|C5|C2] [N] [flags] [lib]   (flags: model.tuning flags, e.g. 256 = blocked forward)"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import qmcnn_b200 as q
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = bench.CONFIGS[name]
N = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["chains"]
flags = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0
if len(sys.argv) > 4:                       # A/B: another build of the library
    from qmcnn_b200 import _lib
    _lib.LIB_PATH = sys.argv[4]
dev = torch.device("cuda", 0)
Ly, Lx = cfg["shape"]
model = q.DCRBM(cfg["k"], cfg["layers"], 2, device=dev, seed=0)
model.set_flat_params(torch.as_tensor(bench.flat_params(cfg, 1234, bench.SCALE)))
model.tuning = dict(flags=flags)
rng = np.random.default_rng(0)
s = torch.as_tensor((rng.integers(0, 2, (N, Ly * Lx)) * 2 - 1).astype(np.int8), device=dev)
w = torch.as_tensor((rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64) / N, device=dev)
work = bench.algorithmic_work(cfg)
fwd_flop = 2.0 * work["mac_full_forward"] if "mac_full_forward" in work else None
# full forward MACs of one sample: sites x sum_l k^2 cin cout
chans = [1] + list(cfg["layers"])
mac = Ly * Lx * sum(cfg["k"] ** 2 * a * b for a, b in zip(chans, chans[1:]))
peak = 74.2e12


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cache = torch.empty(N * model.handle((Ly, Lx)).cache_floats, dtype=torch.float32, device=dev)
t_f = timeit(lambda: model.forward_unpadded(s, (Ly, Lx), want_factors=False, want_logpsi=True, cache=cache))
t_b = timeit(lambda: q.logpsi_gradient(model, s, w, (Ly, Lx)))
print("%s N=%d flags=%d: forward %.3f ms = %.2f TFLOP/s = %.2f of FP32 peak; gradient (forward + backward) %.3f ms, "
      "backward alone %.3f ms = %.2f TFLOP/s = %.2f of peak (2 x forward MACs)"
      % (name, N, flags, t_f, 2 * mac * N / t_f / 1e9, 2 * mac * N / t_f / 1e9 / 74.2, t_b, t_b - t_f,
         4 * mac * N / (t_b - t_f) / 1e9, 4 * mac * N / (t_b - t_f) / 1e9 / 74.2))
