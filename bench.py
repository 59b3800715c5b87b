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
SCALE = 1e-2          # models.py:9,73 (override: --sigma)


def round4(v):
    return (v + 3) & ~3


def cache_mbytes(cfg, S):
    """Per-GPU activation cache (qmc_cache_floats restated: hidden planes padded to 4 channels + fRe + fIm)."""
    n = cfg["shape"][0] * cfg["shape"][1]
    chans = [2 * cfg["alpha"]] if cfg["kind"] == "crbm" else list(cfg["layers"])
    floats = sum(round4(c) * n for c in chans[:-1]) + 2 * round4(n)
    return S * floats * 4 * (8 if cfg.get("sym") else 1) / 1e6


def config_dict(name, cfg, S, world, sigma):
    """`config` of the JSON line - identical for the CUDA arm and the reference arm."""
    mb = cache_mbytes(cfg, S)
    return {"workload": workload_name(name, dict(cfg, chains=S)), "chains_total": S * world, "sigma": sigma,
            "l2": ("activation caches %.0f MB per GPU exceed the 126 MB L2 (inputs larger than L2); every "
                   "step starts from a full forward that rewrites them" if mb > 126 else
                   "activation caches %.0f MB per GPU fit the L2 and are not flushed: parity-test "
                   "configuration, not the headline workload") % mb,
            "step": "mcmc_op + local energies + moment allreduce + gradient + gradient allreduce + Adam"}


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


def flat_params(cfg, seed, sigma=SCALE):
    """Synthetic parameters ~N(0, sigma) in the reference variable order (no oracle import)."""
    rng = np.random.default_rng(seed)
    k = cfg["k"]
    if cfg["kind"] == "crbm":
        shapes = [(k, k, 1, 2 * cfg["alpha"]), (2,), (2 * cfg["alpha"],)]
    else:
        ch = [1] + list(cfg["layers"])
        shapes = []
        for a, b in zip(ch, ch[1:]):
            shapes += [(k, k, a, b), (b,)]
    return np.concatenate([(sigma * rng.standard_normal(s)).astype(np.float32).ravel() for s in shapes])


# ------------------------------------------------------------------------------ reference algorithm on host cores
class CpuReference(object):
    """The reference algorithm (full network per proposal, sampler.py:117-133; window-trick local energy,
    mcmc_tf.py:59-90) restated in torch-CPU (oracle/torch_ref.py; TensorFlow itself is not installable here) on a
    BOUNDED sample of the workload: `chains` chains x `its` Metropolis iterations per step, then the local energy
    of `ne` samples.  ONE code path for `--impl reference` and for the `cpu_baseline` leg of the CUDA arm."""

    NOTE = ("reference algorithm (full network per proposal, sampler.py:117-133) restated in torch-CPU "
            "(oracle/torch_ref.py); TensorFlow itself is not installable here")

    def __init__(self, cfg, sigma, chains=256, its=256, ne=4):
        import oracle
        from oracle import torch_ref
        self.torch_ref, self.cfg = torch_ref, cfg
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        rng = np.random.default_rng(1234)
        if cfg["kind"] == "crbm":
            om = oracle.CRBM(cfg["k"], (cfg["k"] - 1) // 2, cfg["alpha"], 2, rng=rng, scale=sigma)
        else:
            om = oracle.DCRBM(cfg["k"], cfg["layers"], 2, rng=rng, scale=sigma)
        self.tm = torch_ref.TorchModel(om)
        self.S, self.its, self.ne = chains, its, ne
        self.rng = np.random.default_rng(0)
        Ly, Lx = cfg["shape"]
        self.states = torch.as_tensor((self.rng.integers(0, 2, (chains, Ly, Lx)) * 2 - 1).astype(np.float32))

    def step(self):
        """Returns (seconds in the sweep, seconds in the energy)."""
        cfg = self.cfg
        Ly, Lx = cfg["shape"]
        pos = torch.as_tensor(self.rng.integers(0, Ly * Lx, (self.its, self.S, cfg["flips"])).astype(np.int64))
        u = torch.as_tensor(self.rng.random((self.its, self.S)).astype(np.float32))
        t0 = time.perf_counter()
        cur, _ = self.torch_ref.metropolis_steps(self.tm, self.states, pos, u)
        t1 = time.perf_counter()
        if cfg["ham"] == "tfim" and self.ne:
            self.torch_ref.ising_energy(self.tm, cur[:self.ne], H=cfg["H"])
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def sample_text(self):
        return "%d chains x %d full-network Metropolis its per step (of %d chains x %d its), energy on %d samples" % (
            self.S, self.its, self.cfg["chains"], sample_its(self.cfg), self.ne)

    def measure(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        ts = te = 0.0
        for _ in range(steps):
            a, b = self.step()
            ts += a
            te += b
        value = self.S * self.its * steps / ts
        return value, ts, te, {"value": value, "unit": "proposals/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": self.sample_text() + ", %d step(s)" % steps,
                               "local_energies_per_s": (self.ne * steps / te) if te > 0 else None, "note": self.NOTE}


def run_reference(args, cfg, name):
    """`--impl reference`: the reference algorithm on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ref = CpuReference(cfg, args.sigma, args.ref_chains, args.ref_its, args.ref_energy)
    value, ts, te, base = ref.measure(args.steps, args.warmup)
    S = args.chains or cfg["chains"]
    line = {
        "impl": "reference", "metric": "metropolis_proposals_per_s", "value": value, "unit": "proposals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (ts + te) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, cfg, S, world, args.sigma),
        "local_energies_per_s": base["local_energies_per_s"],
        "cpu_baseline": base,
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
ISSUE_PROFILES = {   # config: (ncu summary under profiles/, proposals of the captured launch = 200 sweep its x chains)
    "C2": ("r02_k_sweep_w28_C2_metrics.txt", 200 * 4096),
    "C4": ("r02_k_sweep_sym_C4_metrics.txt", 200 * 8192),
}


def issue_bound(name, proposals_per_s_per_gpu, num_sms, clk, root=None):
    """Instruction-issue roofline of the sweep kernels whose captures are listed in ISSUE_PROFILES, or None.  Never raises:
    the bench line must not depend on a profile file."""
    try:
        if name not in ISSUE_PROFILES:
            return None
        fn, nprop = ISSUE_PROFILES[name]
        inst = None
        for ln in open(os.path.join(root or ROOT, "profiles", fn)):
            f = ln.split()
            if len(f) >= 2 and f[0] == "smsp__inst_executed.sum":
                inst = float(f[1])
                break
        mhz = (clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz")
        if not inst or not mhz:
            return None
        per_prop = inst / nprop
        peak = num_sms * 4 * mhz * 1e6                    # one warp instruction per scheduler and cycle
        return {"warp_instructions_per_proposal_from_profile": per_prop, "profile": "profiles/" + fn,
                "peak_warp_instructions_per_s": peak, "achieved_warp_instructions_per_s": proposals_per_s_per_gpu * per_prop,
                "frac": proposals_per_s_per_gpu * per_prop / peak}
    except Exception:
        return None


def algorithmic_work_total(cfg):
    """Per proposal, incl. the symmetry images and the number of flipped sites (SURVEY 8d closed forms are per
    single-site window of one network)."""
    w = algorithmic_work(cfg)
    mult = (8 if cfg.get("sym") else 1) * cfg["flips"]
    return {k: v * mult for k, v in w.items()}


def run_cuda(args, cfg, name):
    import torch.distributed as dist
    import qmcnn_b200 as q
    from qmcnn_b200 import _lib
    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)

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
    params_host = torch.as_tensor(flat_params(cfg, 1234, args.sigma)).pin_memory()
    model.set_flat_params(params_host)
    if args.tuning:       # experiments only (model.tuning -> qmc_model_desc.reserved): "max_warps=8,ip_group=2,flags=0x40"
        model.tuning = {k: int(v, 0) for k, v in (kv.split("=") for kv in args.tuning.split(","))}
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
    # the timed step IS the product's optimisation step (mcmc_tf.py:218-222): OptimizeStep.run()
    opt = q.optimize_op(sampler, model, energy_fn)
    opt.record_events = True

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        opt.run(new_samples=True)
    sync()
    # kernels of the sweep segment of one step (library's own launch counter; the reset forward is part of mcmc_op)
    lib = _lib.load()
    l0 = lib.qmc_launch_count()
    sampler._sweep(0, its)
    sweep_launches = max(int(lib.qmc_launch_count() - l0) - 1, 1)     # minus the parameter repack of handle()
    if cfg.get("sym"):
        sweep_kernel = "k_sweep_sym"
    elif getattr(sampler, "_nd", False):
        sweep_kernel = "k_nd_sweep"
    elif cfg["kind"] == "dcrbm" and sweep_launches > 1:
        sweep_kernel = "k_sweep_ip (time-sliced: one chunk of one chain per warp slot and launch)"
    else:
        sweep_kernel = "k_sweep_w8 / k_sweep_w16 / k_sweep_ip (single launch)"
    sync()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_start, t_end = ev(), ev()
    sampler._n_accept.zero_()
    launches0 = lib.qmc_launch_count()
    sync()
    t_start.record()
    marks = []
    for _ in range(args.steps):
        last_e = opt.run(new_samples=True)
        marks.append(opt.last_events)
    t_end.record()
    sync()
    clk = clocks.stop() if clocks else None
    gpu_launches = int(lib.qmc_launch_count() - launches0)     # counted by the library itself
    import ctypes
    prof = (ctypes.c_ulonglong * 21)()
    lib.qmc_diag_ip_profile(prof)            # all zero unless the library is a -DQMC_IP_PROFILE=1 variant build
    if prof[8] and rank == 0:
        names = ["draw+top barrier", "spin tile+gathers", "layer 0", "layer barriers", "conv loops", "tanh epilogues",
                 "head", "accept+commit"]
        print("k_sweep_ip phase cycles per proposal per warp: " +
              ", ".join("%s %.1fk" % (nm, prof[i] / prof[8] / 1e3) for i, nm in enumerate(names)) +
              "; total %.1fk" % (sum(prof[:8]) / prof[8] / 1e3) +
              "; task duration by warp index relative to warp 0: " +
              " ".join("%.3f" % (prof[9 + w] / max(prof[9], 1)) for w in range(12)), file=sys.stderr)
    elapsed_ms = t_start.elapsed_time(t_end)
    seg = {"sweep": 0.0, "energy": 0.0, "gradient": 0.0}
    for e0, e1, e2, e3 in marks:
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
    e_mean = float(last_e.real.mean())
    e_err = float(last_e.real.std() / np.sqrt(S))

    # ---- end to end: the SAME step through the public API with HOST buffers: every step copies its inputs
    # (parameters, initial lattices) from pinned host memory and reads its results (energies, updated
    # parameters, samples) back
    host_init = torch.empty((S, n), dtype=torch.int8).pin_memory()
    host_init.copy_(torch.randint(0, 2, (S, n), dtype=torch.int8) * 2 - 1)
    host_samples = torch.empty((S, n), dtype=torch.int8).pin_memory()
    host_e = torch.empty(S, dtype=torch.complex64).pin_memory()
    host_params_out = torch.empty_like(params_host).pin_memory()

    def e2e_step():
        base_model.flat.copy_(params_host, non_blocking=True)                  # H2D parameters
        sampler.feed(initial_states=host_init.to(dev, non_blocking=True))      # H2D lattices
        energies = opt.run(new_samples=True)                                   # sample, E_loc, gradient, allreduces, Adam
        host_samples.copy_(sampler.samples_int8(), non_blocking=True)          # D2H samples
        host_e.copy_(energies, non_blocking=True)                              # D2H energies
        host_params_out.copy_(base_model.flat, non_blocking=True)              # D2H updated parameters
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

    if rank == 0:
        work = algorithmic_work_total(cfg)
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
        props_per_step = float(S) * its
        ach_tflops = props_per_step * work["flop"] / sweep_s * 1e-12
        # which of the three rooflines bounds one proposal:
        # t = max(F / FP32_peak, T / MUFU_peak, B / HBM_peak); ~2 MUFU per tanh, ~5 per complex log-2cosh
        mufu_ops = 2.0 * work["tanh"] + 5.0 * work["logcosh"]
        t_f = work["flop"] / (fp32_peak * 1e12) if fp32_peak else 0.0
        t_m = mufu_ops / (mufu.value * 1e9) if mufu.value else 0.0
        t_b = work["window_bytes"] * (1 + accept_rate) / (hbm_peak * 1e9)
        bound = "fp32_fma" if t_f >= max(t_m, t_b) else ("mufu" if t_m >= t_b else "hbm")
        roofline = {
            "kernel": sweep_kernel, "bound": bound, "bound_times_ns": {"fp32": t_f * 1e9, "mufu": t_m * 1e9, "hbm": t_b * 1e9},
            "frac_of_bound": max(t_f, t_m, t_b) / (sweep_s / props_per_step),
            "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": ach_tflops / fp32_peak if fp32_peak else None,
            "ffma_tflops": f32.value, "ffma2_tflops": f32x2.value,
            "peak_source": "own FFMA / FFMA2 microbenchmark (qmc_diag_peaks2) on this GPU, max of the two; MEASURED_PEAKS.json has no FP32 peak",
            "algorithmic_flop_per_proposal": work["flop"],
            "hbm": {"achieved": props_per_step * work["window_bytes"] * (1 + accept_rate) / sweep_s * 1e-9,
                    "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "algorithmic_bytes_per_proposal": work["window_bytes"] * (1 + accept_rate)},
            "mufu_gops_peak": mufu.value,
            "note": "'achieved' = algorithmic FLOP of the sweep / sweep time (live CUDA events); 'frac' is against the FP32 "
                    "FMA peak of the CUDA cores, 'frac_of_bound' against whichever of FP32 / MUFU / HBM bounds one proposal "
                    "(SURVEY.md section 8d).  Tensor cores: measured no-go for these 16-channel layers "
                    "(profiles/r02_tc_layer_proto.txt)",
            "traffic": None,
        }
        roofline["sweep_kernel_launches_per_step"] = sweep_launches
        roofline["proposals_per_launch"] = props_per_step / max(sweep_launches, 1)
        # DRAM traffic is NOT measured by this run: it is the ncu figure of the committed profile x this run's launch size
        for prof_name in ("r02_sweep_traffic.json", "r01_sweep_traffic.json"):
            prof = os.path.join(ROOT, "profiles", prof_name)
            if name in ("C3", "C5") and os.path.exists(prof):
                try:
                    roofline["traffic"] = json.load(open(prof))["dram_bytes_per_proposal"] * roofline["proposals_per_launch"]
                    roofline["traffic_from_profile"] = ("profiles/%s: ncu dram__bytes_read+write per proposal of the C3 capture "
                                                        "x proposals per launch of this run (not re-measured here)" % prof_name)
                except Exception:
                    pass
                break
        # C2 (small model, classic kernel) and C4 (symmetry sweep) are bound by instruction issue, not by FP32 / MUFU / HBM
        # (profiles/r02_summary.md, "Classic kernels"): warp instructions per proposal of the committed ncu capture (NOT
        # re-measured here) against what the SMs' four schedulers can issue at the clock this run saw
        try:
            roofline["issue"] = issue_bound(name, proposals / (sw_ms * 1e-3) / world,
                                            torch.cuda.get_device_properties(dev).multi_processor_count, clk)
        except Exception:
            roofline["issue"] = None
        line = {
            "metric": "metropolis_proposals_per_s", "value": value, "unit": "proposals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(name, cfg, S, world, args.sigma),
            "local_energies_per_s": float(S) * args.steps * world / (en_ms * 1e-3),
            "sweep_proposals_per_s": proposals / (sw_ms * 1e-3),
            "segments_ms_per_step": {"sweep": sw_ms / args.steps, "energy": en_ms / args.steps,
                                     "gradient": gr_ms / args.steps},
            "acceptance_rate": accept_rate, "energy_per_spin": [e_mean, e_err],
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "proposals/s",
                    "h2d_bytes_per_step": S * n + params_host.numel() * 4,
                    "d2h_bytes_per_step": S * n + S * 8 + params_host.numel() * 4, "steps": e2e_steps,
                    "step": "H2D parameters + lattices, OptimizeStep.run() (the timed step), D2H samples + energies + parameters"},
            "gpu_launches": gpu_launches,       # this library's kernels in the timed region (qmc_launch_count)
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample (~10-30 s) of the same workload on the host cores, same code path as --impl reference
            line["cpu_baseline"] = CpuReference(cfg, args.sigma, args.ref_chains, args.ref_its, args.ref_energy).measure(1, 1)[3]
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
    ap.add_argument("--lib", default="", help="A/B measurements: load this build of the library instead of the in-tree one")
    ap.add_argument("--tuning", default="", help="experiments: model.tuning as key=int pairs, e.g. max_warps=8,ip_group=2")
    ap.add_argument("--sigma", type=float, default=SCALE,
                    help="std of the synthetic parameters (models.py SCALE = 1e-2: acceptance ~ 1; 1e-1: non-trivial)")
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
