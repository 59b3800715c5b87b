#!/usr/bin/env python
"""bench.py - Metropolis proposals/s and local energies/s of the VMC hot path.

  python bench.py --gpus N --steps K --warmup W           # the CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One "step" is one pass of the hot path over one batch: a full `mcmc_op`
(reset forward + sample_its Metropolis iterations of every chain), the local
energies of the drawn samples, the energy-moment allreduce, the log-psi
gradient, the gradient allreduce and the TF-1 Adam update - i.e. one
`sess.run(optimize)` of mcmc_tf.py:218-222.  Workload (BASELINE.json configs[2],
the 20x20 configuration the metric is quoted on, chains fixed per GPU = weak
scaling): 20x20 TFIM, DCRBM(k=3, [16,16,16,16,16,8]), 4096 chains per GPU
(32768 at 8 GPUs), sample_its = 16001, synthetic random-init parameters ~N(0, 1e-2)
and iid +-1 initial lattices.

The JSON line: `value` = proposals of all ranks / device time of the K steps
(CUDA events, barrier + sync both sides, max over ranks); `e2e` = the same
through the public API with host buffers (pinned H2D of the initial lattices and
parameters, D2H of samples and energies inside the timed region); `roofline`
for the dominant kernel (k_sweep); `cpu_baseline` = the reference algorithm
(full network per proposal, torch-CPU restatement in oracle/torch_ref.py) on the
host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

CONFIGS = {
    # name: (shape, model kind, model kwargs, chains per GPU, H, hamiltonian, num_flips)
    "C1": dict(shape=(6, 6), kind="crbm", k=5, alpha=4, chains=64, H=1.0, ham="tfim", flips=1),
    "C2": dict(shape=(10, 10), kind="dcrbm", k=3, layers=[8, 8, 8], chains=4096, H=3.0, ham="tfim", flips=1),
    "C3": dict(shape=(20, 20), kind="dcrbm", k=3, layers=[16, 16, 16, 16, 16, 8], chains=4096, H=1.0,
               ham="tfim", flips=1),
    # tuning experiment (not a BASELINE config): 4-layer variant whose tiles allow 14+ warps per SM
    "X4": dict(shape=(20, 20), kind="dcrbm", k=3, layers=[16, 16, 16, 8], chains=4096, H=1.0, ham="tfim", flips=1),
    "C4": dict(shape=(10, 10), kind="crbm", k=5, alpha=4, chains=8192, H=1.0, ham="heis", flips=2, sym=True),
    "C5": dict(shape=(40, 40), kind="dcrbm", k=3, layers=[16, 16, 16, 16, 16, 8], chains=8192, H=1.0,
               ham="tfim", flips=1),
}
SCALE = 1e-2          # models.py:9,73


def algorithmic_work(cfg):
    """SURVEY.md section 8(d) closed forms, per proposal (incremental algorithm)."""
    k = cfg["k"]
    Ly = cfg["shape"][0]
    chans = [2 * cfg["alpha"]] if cfg["kind"] == "crbm" else list(cfg["layers"])
    D = len(chans)
    w = [min(l * (k - 1) + 1, Ly) for l in range(1, D + 1)]
    mac = w[0] ** 2 * chans[0]
    for l in range(1, D):
        mac += w[l] ** 2 * k * k * chans[l - 1] * chans[l]
    tanh = sum(w[l] ** 2 * chans[l] for l in range(D - 1))
    logcosh = w[D - 1] ** 2 * (chans[D - 1] // 2)
    bytes_rw = sum(w[l] ** 2 * chans[l] * 4 for l in range(D))
    return dict(flop=2.0 * mac, mac=mac, tanh=tanh, logcosh=logcosh, window_bytes=bytes_rw)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4)
                          if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def build_models(cfg, seed):
    import oracle
    rng = np.random.default_rng(seed)
    if cfg["kind"] == "crbm":
        om = oracle.CRBM(cfg["k"], (cfg["k"] - 1) // 2, cfg["alpha"], 2, rng=rng, scale=SCALE)
    else:
        om = oracle.DCRBM(cfg["k"], cfg["layers"], 2, rng=rng, scale=SCALE)
    return om


def flat_params(cfg, seed):
    """Synthetic parameters ~N(0, SCALE) in the reference variable order (no oracle import)."""
    rng = np.random.default_rng(seed)
    k = cfg["k"]
    if cfg["kind"] == "crbm":
        shapes = [(k, k, 1, 2 * cfg["alpha"]), (2,), (2 * cfg["alpha"],)]
    else:
        ch = [1] + list(cfg["layers"])
        shapes = []
        for a, b in zip(ch, ch[1:]):
            shapes += [(k, k, a, b), (b,)]
    return np.concatenate([(SCALE * rng.standard_normal(s)).astype(np.float32).ravel() for s in shapes])


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, cfg, name):
    """The reference algorithm on the host cores (torch-CPU restatement; TF cannot be installed).
    Each step: a bounded sample of the workload - `ref_chains` chains x `ref_its` full-network
    Metropolis iterations, then the window-trick local energy of `ref_energy` samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    om = build_models(cfg, 1234)
    tm = torch_ref.TorchModel(om)
    Ly, Lx = cfg["shape"]
    S, its, ne = args.ref_chains, args.ref_its, args.ref_energy
    rng = np.random.default_rng(0)
    states = torch.as_tensor((rng.integers(0, 2, (S, Ly, Lx)) * 2 - 1).astype(np.float32))

    def step():
        pos = torch.as_tensor(rng.integers(0, Ly * Lx, (its, S, cfg["flips"])).astype(np.int64))
        u = torch.as_tensor(rng.random((its, S)).astype(np.float32))
        t0 = time.perf_counter()
        cur, _ = torch_ref.metropolis_steps(tm, states, pos, u)
        t1 = time.perf_counter()
        if cfg["ham"] == "tfim":
            torch_ref.ising_energy(tm, cur[:ne], H=cfg["H"])
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    for _ in range(args.warmup):
        step()
    ts, te = 0.0, 0.0
    for _ in range(args.steps):
        a, b = step()
        ts += a
        te += b
    props = S * its * args.steps
    value = props / ts
    sample = "%d chains x %d full-network Metropolis its per step (of %d chains x %d its), energy on %d samples" % (
        S, its, cfg["chains"], sample_its(cfg), ne)
    line = {
        "impl": "reference", "metric": "metropolis_proposals_per_s", "value": value, "unit": "proposals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (ts + te) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(name, cfg), "sample": sample},
        "local_energies_per_s": (ne * args.steps / te) if te > 0 else None,
        "cpu_baseline": {"value": value, "unit": "proposals/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample,
                         "note": "reference algorithm restated in torch-CPU (oracle/torch_ref.py); "
                                 "TensorFlow itself is not installable here"},
        "e2e": {"value": value, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def sample_its(cfg):
    n = cfg["shape"][0] * cfg["shape"][1]
    return 4 * 10 * n + 1          # sampler.py:29-35 with samples_per_sampler = 1


def workload_name(name, cfg):
    m = ("CRBM(k=%d,alpha=%d)" % (cfg["k"], cfg["alpha"]) if cfg["kind"] == "crbm"
         else "DCRBM(k=%d,layers=%s)" % (cfg["k"], cfg["layers"]))
    if cfg.get("sym"):
        m += " + D4xT symmetry average"
    return "%s: %dx%d %s h=%g, %s, %d chains/GPU, sample_its=%d" % (
        name, cfg["shape"][0], cfg["shape"][1], cfg["ham"].upper(), cfg["H"], m, cfg["chains"], sample_its(cfg))


# ------------------------------------------------------------------------------ CUDA arm
def cpu_baseline(cfg, budget_s=20.0):
    """Bounded CPU sample of the same workload, rank 0, N=1 only."""
    import oracle
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    om = build_models(cfg, 1234)
    tm = torch_ref.TorchModel(om)
    Ly, Lx = cfg["shape"]
    S = 256
    rng = np.random.default_rng(0)
    states = torch.as_tensor((rng.integers(0, 2, (S, Ly, Lx)) * 2 - 1).astype(np.float32))
    its = 64
    pos = torch.as_tensor(rng.integers(0, Ly * Lx, (its, S, cfg["flips"])).astype(np.int64))
    u = torch.as_tensor(rng.random((its, S)).astype(np.float32))
    torch_ref.metropolis_steps(tm, states, pos[:4], u[:4])          # warm-up
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s * 0.7:
        torch_ref.metropolis_steps(tm, states, pos, u)
        done += its * S
    ts = time.perf_counter() - t0
    ne, te, nd = 4, 0.0, 0
    if cfg["ham"] == "tfim":
        t1 = time.perf_counter()
        while time.perf_counter() - t1 < budget_s * 0.3:
            torch_ref.ising_energy(tm, states[:ne], H=cfg["H"])
            nd += ne
        te = time.perf_counter() - t1
    return {"value": done / ts, "unit": "proposals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d chains x %d full-network Metropolis its repeated for %.0f s; energy on %d samples for %.0f s"
                      % (S, its, ts, ne, te),
            "local_energies_per_s": (nd / te) if te > 0 else None,
            "note": "reference algorithm (full network per proposal, sampler.py:117-133) restated in torch-CPU; "
                    "TensorFlow itself is not installable here"}


def run_cuda(args, cfg, name):
    import torch.distributed as dist
    import qmcnn_b200 as q
    from qmcnn_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Ly, Lx = cfg["shape"]
    n = Ly * Lx
    S = args.chains or cfg["chains"]

    class BenchSampler(q.Sampler):
        MAX_NUM_SAMPLERS = 10 ** 9        # the reference cap must be lifted for C2-C5

    if cfg["kind"] == "crbm":
        model = q.CRBM(cfg["k"], (cfg["k"] - 1) // 2, cfg["alpha"], 2, device=dev, seed=0)
    else:
        model = q.DCRBM(cfg["k"], cfg["layers"], 2, device=dev, seed=0)
    params_host = torch.as_tensor(flat_params(cfg, 1234)).pin_memory()
    model.set_flat_params(params_host)
    base_model = model
    if cfg.get("sym"):
        model = q.SymmetrizedModel(model)      # D4 x| T averaged amplitude (8 filter images)
    sampler = BenchSampler(model, (Ly, Lx), model.r, S, cfg["flips"], seed=1234, chain_id0=rank * S)
    if args.sweep_its:
        sampler.sample_its = args.sweep_its
        sampler.therm_its = args.sweep_its - 1
    its = sampler.sample_its
    if cfg["ham"] == "tfim":
        energy_fn = lambda s: q.ising_energy(model, s, system_shape=(Ly, Lx), H=cfg["H"])
    else:
        energy_fn = lambda s: q.heisenberg_energy(model, s, system_shape=(Ly, Lx))
    opt = q.optimize_op(sampler, model, energy_fn)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    seg = {"sweep": 0.0, "energy": 0.0, "gradient": 0.0}

    def step(timed):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        sampler.mcmc_op()
        e1.record()
        samples = sampler.samples_int8()
        energies = energy_fn(samples)
        mom = torch.stack([torch.tensor(float(S), device=dev, dtype=torch.float64),
                           energies.real.double().sum(), energies.imag.double().sum()])
        if world > 1:
            dist.all_reduce(mom)
        e2.record()
        e_mean = torch.complex(mom[1] / mom[0], mom[2] / mom[0]).to(torch.complex64)
        grad = q.logpsi_gradient(model, samples, (energies - e_mean) / mom[0].float(), (Ly, Lx))
        if world > 1:
            dist.all_reduce(grad)
        opt.optimizer.step(grad)
        e3.record()
        return (e0, e1, e2, e3), energies

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    sync()
    # launches of the sweep kernel in one mcmc_op (the reset forward is one more launch)
    l0 = _lib.load().qmc_launch_count()
    sampler._sweep(0, its)
    sweep_launches = int(_lib.load().qmc_launch_count() - l0)
    # the probe counts the parameter repack + the sweep launches: the classic persistent kernel is always ONE
    # launch; more than one means the in-place kernel, time-sliced
    sweep_launches = max(sweep_launches - 1, 1)
    sweep_kernel = "k_sweep_ip" if sweep_launches > 1 else "k_sweep / k_sweep_ip (single launch: chains <= warp slots)"
    sync()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    t_start, t_end = ev(), ev()
    sampler._n_accept.zero_()
    launches0 = _lib.load().qmc_launch_count()
    sync()
    t_start.record()
    marks = [step(True) for _ in range(args.steps)]
    t_end.record()
    sync()
    clk = clocks.stop() if clocks else None
    gpu_launches = int(_lib.load().qmc_launch_count() - launches0)     # counted by the library itself
    elapsed_ms = t_start.elapsed_time(t_end)
    for (e0, e1, e2, e3), _ in marks:
        seg["sweep"] += e0.elapsed_time(e1)
        seg["energy"] += e1.elapsed_time(e2)
        seg["gradient"] += e2.elapsed_time(e3)
    tmax = torch.tensor([elapsed_ms, seg["sweep"], seg["energy"], seg["gradient"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms, sw_ms, en_ms, gr_ms = [float(v) for v in tmax.cpu()]
    proposals = float(S) * its * args.steps * world
    value = proposals / (elapsed_ms * 1e-3)
    accept_rate = sampler.acceptance_count / (float(S) * its * args.steps)
    last_e = marks[-1][1]
    e_mean = float(last_e.real.mean())
    e_err = float(last_e.real.std() / np.sqrt(S))

    # ---- end to end through the public API with host buffers (rank-local, all ranks run it)
    host_init = torch.empty((S, n), dtype=torch.int8).pin_memory()
    host_init.copy_(torch.randint(0, 2, (S, n), dtype=torch.int8) * 2 - 1)
    host_samples = torch.empty((S, n), dtype=torch.int8).pin_memory()
    host_e = torch.empty(S, dtype=torch.complex64).pin_memory()

    def e2e_step():
        base_model.flat.copy_(params_host, non_blocking=True)                  # H2D parameters
        sampler.feed(initial_states=host_init.to(dev, non_blocking=True))      # H2D lattices
        sampler.new_samples = True
        sampler.mcmc_op()
        smp = sampler.samples_int8()
        host_samples.copy_(smp, non_blocking=True)                             # D2H samples
        host_e.copy_(energy_fn(smp), non_blocking=True)                        # D2H energies
        torch.cuda.current_stream().synchronize()
    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = float(S) * its * e2e_steps * world / float(e2e_s.item())

    cache_mb = S * sampler._h.cache_floats * 4 * (8 if cfg.get("sym") else 1) / 1e6
    if rank == 0:
        work = algorithmic_work(cfg)
        lib = _lib.load()
        import ctypes
        f32, f32x2, mufu = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        lib.qmc_diag_peaks2(local_rank, ctypes.byref(f32), ctypes.byref(f32x2), ctypes.byref(mufu))
        fp32_peak = max(f32.value, f32x2.value)     # the denominator is the better of FFMA and FFMA2
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # the sweep segment of one step: the reset forward (<1%) + the launches of the sweep kernel.  For deep
        # models that is k_sweep_ip, time-sliced by the host into full-wave launches (one chunk of one chain per
        # warp slot); total proposals / total time equals the launch-weighted mean of per-launch rates.
        sweep_s = sw_ms * 1e-3 / args.steps
        props_per_launch = float(S) * its
        ach_tflops = props_per_launch * work["flop"] / sweep_s * 1e-12
        # which of the three rooflines bounds one proposal:
        # t = max(F / FP32_peak, T / MUFU_peak, B / HBM_peak); ~2 MUFU per tanh, ~5 per complex log-2cosh
        mufu_ops = 2.0 * work["tanh"] + 5.0 * work["logcosh"]
        t_f = work["flop"] / (fp32_peak * 1e12) if fp32_peak else 0.0
        t_m = mufu_ops / (mufu.value * 1e9) if mufu.value else 0.0
        t_b = work["window_bytes"] * (1 + accept_rate) / (hbm_peak * 1e9)
        bound = "fp32_fma" if t_f >= max(t_m, t_b) else ("mufu" if t_m >= t_b else "hbm")
        roofline = {
            "kernel": sweep_kernel, "bound": bound, "bound_times_ns": {"fp32": t_f * 1e9, "mufu": t_m * 1e9, "hbm": t_b * 1e9},
            "frac_of_bound": max(t_f, t_m, t_b) / (sweep_s / props_per_launch),
            "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": ach_tflops / fp32_peak if fp32_peak else None,
            "ffma_tflops": f32.value, "ffma2_tflops": f32x2.value,
            "peak_source": "own FFMA / FFMA2 microbenchmark (qmc_diag_peaks2) on this GPU, max of the two; MEASURED_PEAKS.json has no FP32 peak",
            "algorithmic_flop_per_proposal": work["flop"],
            "hbm": {"achieved": props_per_launch * work["window_bytes"] * (1 + accept_rate) / sweep_s * 1e-9,
                    "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "algorithmic_bytes_per_proposal": work["window_bytes"] * (1 + accept_rate)},
            "mufu_gops_peak": mufu.value,
            "note": "bound is the FP32 FMA pipe of the CUDA cores (SURVEY.md section 8d: C <= 16 channels and fp32 "
                    "parity rule out tensor cores; the per-proposal window is served from L2/HBM at ~0.14 of the HBM "
                    "roofline, reported under 'hbm'); 'achieved' = algorithmic FLOP / sweep time",
            "traffic": None,
        }
        roofline["sweep_kernel_launches_per_step"] = sweep_launches
        roofline["proposals_per_launch"] = props_per_launch / max(sweep_launches, 1)
        prof = os.path.join(ROOT, "profiles", "r01_sweep_traffic.json")
        if os.path.exists(prof):
            try:
                roofline["traffic"] = json.load(open(prof))["dram_bytes_per_proposal"] * roofline["proposals_per_launch"]
                roofline["traffic_source"] = "profiles/r01_sweep_traffic.json (ncu dram__bytes per proposal x proposals per launch)"
            except Exception:
                pass
        line = {
            "metric": "metropolis_proposals_per_s", "value": value, "unit": "proposals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(name, dict(cfg, chains=S)), "chains_total": S * world,
                       "l2": ("activation caches %.0f MB per GPU exceed the 126 MB L2 (inputs larger than L2); every "
                              "step starts from a full forward that rewrites them" if cache_mb > 126 else
                              "activation caches %.0f MB per GPU fit the L2 and are not flushed: parity-test "
                              "configuration, not the headline workload") % cache_mb,
                       "step": "mcmc_op + local energies + moment allreduce + gradient + gradient allreduce + Adam"},
            "local_energies_per_s": float(S) * args.steps * world / (en_ms * 1e-3),
            "sweep_proposals_per_s": proposals / (sw_ms * 1e-3),
            "segments_ms_per_step": {"sweep": sw_ms / args.steps, "energy": en_ms / args.steps,
                                     "gradient": gr_ms / args.steps},
            "acceptance_rate": accept_rate, "energy_per_spin": [e_mean, e_err],
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "proposals/s",
                    "h2d_bytes_per_step": S * n + params_host.numel() * 4,
                    "d2h_bytes_per_step": S * n + S * 8, "steps": e2e_steps},
            "gpu_launches": gpu_launches,       # this library's kernels in the timed region (qmc_launch_count)
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the config's)")
    ap.add_argument("--sweep-its", type=int, default=0, help="Metropolis iterations per step (default sample_its)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-chains", type=int, default=256)
    ap.add_argument("--ref-its", type=int, default=256)
    ap.add_argument("--ref-energy", type=int, default=4)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, args.config)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the qmcnn_b200 path has no CPU fallback "
                             "(use --impl reference for the host-core arm)")
        run_cuda(args, cfg, args.config)


if __name__ == "__main__":
    main()
