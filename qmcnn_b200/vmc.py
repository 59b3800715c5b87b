"""The reference's optimisation loop (module-level script of ``mcmc_tf.py:197-236``) as a
function and a command line: build the model and the two samplers, then per iteration run
``optimize_op`` (sample -> local energies -> gradient of ``loss_op`` -> TF-1 Adam), keep the
chains across iterations (``sampler.new_samples = it == 0``, ``mcmc_tf.py:219-221``), print
``It %d, E=%.5f (%.2e) %.1fs`` and, every ``eval_freq`` iterations, the energy of a fresh
``eval_sampler`` run evaluated in batches (``mcmc_tf.py:228-234``).

    python -m qmcnn_b200.vmc --hamiltonian heisenberg --shape 10 10 --its 200      # the shipped script
    python -m qmcnn_b200.vmc --hamiltonian tfim --h 3.0 --model dcrbm --layers 8 8 8 --k 3 --shape 10 10

Everything numerical happens in the CUDA library; this file is host orchestration only.
"""
from __future__ import division, print_function

import argparse
import sys
from functools import partial
from time import time

import numpy as np
import torch

from . import mcmc
from .mcmc import batched_op, heisenberg_energy, ising_energy, optimize_op
from .models import CRBM, DCRBM
from .sampler import Sampler

# defaults of mcmc_tf.py:14-32
LEARNING_RATE = 3E-3
K = 5
ALPHA = 4
SYSTEM_SHAPE = (10, 10)
H = 1.0
NUM_SAMPLES = 100
OPTIMIZATION_ITS = 10000
ENERGY_BATCH_SIZE = 1000
NUM_EVAL_SAMPLES = 1000
EVAL_FREQ = 20


def make_samplers(model, system_shape, num_samples, num_eval_samples, num_flips, seed=0, group=None,
                  sampler_cls=Sampler):
    """The two samplers of ``mcmc_tf.py:208-209``.  Data parallelism: ``num_samples`` / ``num_eval_samples`` are
    PER RANK and rank g owns the global chains [g * num_samplers, (g + 1) * num_samplers).  Initial lattices and
    proposals are Philox streams keyed by the global chain id, so the ranks draw different chains and the
    all-reduced batch really is world_size times larger (and equals a single-rank run of all the chains)."""
    from . import distributed as D
    rank, _ = D.world(group)
    chains = lambda n: min(n, sampler_cls.MAX_NUM_SAMPLERS)
    sampler = sampler_cls(model, system_shape, model.r, num_samples, num_flips, seed=seed,
                          chain_id0=rank * chains(num_samples))
    eval_sampler = sampler_cls(model, system_shape, model.r, num_eval_samples, num_flips, seed=seed + 1,
                               chain_id0=rank * chains(num_eval_samples))
    return sampler, eval_sampler


def run_vmc(model, system_shape, hamiltonian="heisenberg", h=H, num_samples=NUM_SAMPLES,
            num_eval_samples=NUM_EVAL_SAMPLES, optimization_its=OPTIMIZATION_ITS, eval_freq=EVAL_FREQ,
            learning_rate=LEARNING_RATE, energy_batch_size=ENERGY_BATCH_SIZE, seed=0, group=None,
            log=print, sampler_cls=Sampler):
    """``mcmc_tf.py:201-236``.  Returns a list of dicts, one per iteration
    (``it, energy, stderr, seconds`` and, on evaluation iterations, ``eval_energy, eval_stderr``)."""
    system_shape = tuple(system_shape)
    if hamiltonian == "tfim":
        energy_fn = partial(ising_energy, model, system_shape=system_shape, H=h)            # :203
        num_flips = 1
    elif hamiltonian == "heisenberg":
        energy_fn = partial(heisenberg_energy, model, system_shape=system_shape)            # :205
        num_flips = 2
    else:
        raise ValueError("hamiltonian must be 'tfim' or 'heisenberg'")
    sampler, eval_sampler = make_samplers(model, system_shape, num_samples, num_eval_samples, num_flips, seed,
                                          group, sampler_cls)                                           # :208-209
    optimize = optimize_op(sampler, model, energy_fn, learning_rate=learning_rate, group=group)         # :211
    history = []
    for it in range(optimization_its):                                                      # :218
        start = time()
        e = optimize.run(new_samples=(it == 0)).real                                        # :220-222
        rec = {"it": it + 1, "energy": float(e.mean()), "stderr": float(e.std() / np.sqrt(e.numel())),
               "seconds": time() - start}
        if log:
            log("It %d, E=%.5f (%.2e) %.1fs" % (rec["it"], rec["energy"], rec["stderr"], rec["seconds"]))
        if it > 0 and (it + 1) % eval_freq == 0:                                            # :228
            start = time()
            samples = eval_sampler.mcmc_op()
            bs = energy_batch_size if samples.shape[0] % energy_batch_size == 0 else samples.shape[0]
            ee = batched_op(energy_fn, samples, bs).real                                    # :212-213, 230
            rec["eval_energy"] = float(ee.mean())
            rec["eval_stderr"] = float(ee.std() / np.sqrt(ee.numel()))
            if log:
                log("Eval It %d, E=%.5f (%.2e) %.1fs" % (rec["it"], rec["eval_energy"], rec["eval_stderr"],
                                                        time() - start))
        history.append(rec)
    return history


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--hamiltonian", choices=["tfim", "heisenberg"], default="heisenberg")
    ap.add_argument("--h", type=float, default=H, help="transverse field (TFIM)")
    ap.add_argument("--shape", type=int, nargs=2, default=list(SYSTEM_SHAPE))
    ap.add_argument("--model", choices=["crbm", "dcrbm"], default="crbm")
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--alpha", type=int, default=ALPHA)
    ap.add_argument("--layers", type=int, nargs="+", default=[8, 8, 8])
    ap.add_argument("--samples", type=int, default=NUM_SAMPLES)
    ap.add_argument("--eval-samples", type=int, default=NUM_EVAL_SAMPLES)
    ap.add_argument("--its", type=int, default=OPTIMIZATION_ITS)
    ap.add_argument("--eval-freq", type=int, default=EVAL_FREQ)
    ap.add_argument("--lr", type=float, default=LEARNING_RATE)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    if a.model == "crbm":
        model = CRBM(a.k, (a.k - 1) // 2, a.alpha, 2, seed=a.seed)                          # :202
    else:
        model = DCRBM(a.k, a.layers, 2, seed=a.seed)
    mcmc.SYSTEM_SHAPE, mcmc.K, mcmc.H = tuple(a.shape), model.r, a.h
    run_vmc(model, a.shape, a.hamiltonian, a.h, a.samples, a.eval_samples, a.its, a.eval_freq, a.lr,
            seed=a.seed)
    return 0


if __name__ == "__main__":
    sys.exit(main())
