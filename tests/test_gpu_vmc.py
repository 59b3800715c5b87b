"""End-to-end statistical checks of the sampler and of the optimisation loop on a 4x4 lattice,
where all 65 536 configurations can be enumerated (north_star: "sampled energy means must agree
within 3 sigma"; SURVEY.md section 8f rank 1: the VMC driver loop of mcmc_tf.py:197-236)."""
import itertools

import numpy as np
import pytest
import torch

import oracle
import qmcnn_b200 as q

pytestmark = pytest.mark.gpu

L = 4
SHAPE = (L, L)


def all_states():
    return np.array(list(itertools.product([-1, 1], repeat=L * L)), dtype=np.int32)


def exact_energies(om, ham, h, classes):
    """<E_loc> under |psi|^2 restricted to each class of configurations, by enumeration (float64 oracle)."""
    s = all_states()
    logpsi = om.log_psi(oracle.pad(s.reshape(-1, L, L), SHAPE, [(om.r - 1) // 2] * 2))
    wts = np.exp(2 * (logpsi.real - logpsi.real.max()))
    e = oracle.ising_energy(om, s, SHAPE, om.r, H=h) if ham == "tfim" else oracle.heisenberg_energy(om, s, SHAPE, om.r)
    cls = classes(s)
    return {c: float((wts[cls == c] * e.real[cls == c]).sum() / wts[cls == c].sum()) for c in np.unique(cls)}


@pytest.mark.parametrize("ham,flips", [("tfim", 1), ("heisenberg", 2)])
def test_sampled_energy_mean_within_3_sigma_of_exact(ham, flips):
    """Philox-driven Metropolis chains of the CUDA sampler draw from |psi|^2: the sampled mean of
    E_loc agrees with the enumerated expectation within 3 standard errors.

    The reference's two-flip proposal (two independent uniform sites, sampler.py:95-115) changes the
    number of up spins by -2, 0 or +2, so it conserves that number's parity: a chain stays in the
    parity class of its initial lattice and samples |psi|^2 restricted to it (a property of the
    reference's sampler, reproduced here).  The check is therefore made per class."""
    rng = np.random.default_rng(77)
    om = oracle.CRBM(3, 1, 2, 2, rng=rng, scale=0.3, dtype=np.float64)
    gm = q.CRBM(3, 1, 2, 2, seed=0)
    gm.set_flat_params(om.flat_params().astype(np.float32))
    S = 8192
    classes = (lambda s: np.zeros(len(s), np.int64)) if flips == 1 else (lambda s: ((s > 0).sum(1) // 1) % 2)
    want = exact_energies(om, ham, 1.0, classes)
    GS = type("GS", (q.Sampler,), dict(MAX_NUM_SAMPLERS=S))
    smp = GS(gm, SHAPE, 3, 4 * S, flips, seed=11)               # 8192 chains x 4 samples, 160 steps apart
    init = (rng.integers(0, 2, (S, L * L)) * 2 - 1).astype(np.int32)
    smp.feed(initial_states=init)
    samples = smp.mcmc_op()
    e = (q.ising_energy(gm, samples, system_shape=SHAPE, H=1.0) if ham == "tfim"
         else q.heisenberg_energy(gm, samples, system_shape=SHAPE)).real.double().cpu().numpy()
    # chains are independent; average within a chain first so that residual autocorrelation cannot shrink the error bar
    per_chain = e.reshape(4, S).mean(0)
    chain_cls = classes(init)
    assert np.array_equal(classes(samples.cpu().numpy()), np.tile(chain_cls, 4))       # the class is conserved
    for c, exact in want.items():
        pc = per_chain[chain_cls == c]
        assert pc.size > 1000
        mean, err = pc.mean(), pc.std(ddof=1) / np.sqrt(pc.size)
        assert abs(mean - exact) <= 3 * err, "class %d: sampled %.6f +- %.6f vs exact %.6f" % (c, mean, err, exact)
        assert err < 1e-2 * max(1.0, abs(exact))               # the test has resolving power


def test_vmc_loop_converges_to_exact_ground_state():
    """run_vmc (mcmc_tf.py:197-236): 4x4 TFIM at h = 1 with CRBM(3, alpha 4) approaches the exact
    ground-state energy per spin (sparse diagonalisation) and never goes below it beyond noise."""
    from scipy.sparse import lil_matrix
    from scipy.sparse.linalg import eigsh
    n = L * L
    s = all_states()
    idx = {tuple(r): i for i, r in enumerate(s)}
    Hm = lil_matrix((2 ** n, 2 ** n))
    bonds = oracle.interactions(s, SHAPE).sum((1, 2))
    for i, r in enumerate(s):
        Hm[i, i] = -bonds[i]
        for k in range(n):
            t = r.copy(); t[k] = -t[k]
            Hm[i, idx[tuple(t)]] = -1.0
    e0 = eigsh(Hm.tocsr(), k=1, which="SA")[0][0] / n
    model = q.CRBM(3, 1, 4, 2, seed=3)
    lines = []
    hist = q.run_vmc(model, SHAPE, "tfim", 1.0, num_samples=1000, num_eval_samples=4000, optimization_its=300,
                     eval_freq=50, learning_rate=1e-2, energy_batch_size=1000, seed=5, log=lines.append)
    assert len(hist) == 300 and lines[0].startswith("It 1, E=")
    evals = [r for r in hist if "eval_energy" in r]
    assert [r["it"] for r in evals] == [50, 100, 150, 200, 250, 300]
    first, last = hist[0]["energy"], evals[-1]["eval_energy"]
    assert last < first - 0.3                                   # it learns
    assert last >= e0 - 5 * evals[-1]["eval_stderr"] - 1e-3     # variational bound
    assert last <= e0 + 0.02 * abs(e0), "E = %.5f vs exact %.5f" % (last, e0)
