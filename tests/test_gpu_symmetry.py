"""GPU parity of the symmetry-averaged amplitude (BASELINE config 4: Heisenberg, CRBM(5,2,4,2),
D4 x| T average) against the oracle's brute-force definitions."""
import numpy as np
import pytest
import torch

import oracle
from oracle import symmetry as osym

pytestmark = pytest.mark.gpu


def _pair(kind, L, scale, seed, **kw):
    from gpu_util import make_pair
    import qmcnn_b200 as q
    gm, om = make_pair(kind, L, scale, seed, dtype=np.float64, **kw)
    return q, q.SymmetrizedModel(gm), om


def test_tap_permutations_match_oracle_filter_images():
    import qmcnn_b200 as q
    rng = np.random.default_rng(0)
    for k in (3, 5):
        W = rng.standard_normal((k, k, 2, 3))
        perms = q.symmetry.d4_tap_permutations(k)
        for p, img in enumerate(osym.d4_filter_images(W)):
            assert np.array_equal(W.reshape(k * k, 2, 3)[perms[p]].reshape(W.shape), img)
    # the notebook's checks on the product's own group module
    M = 4
    G = q.symmetry.group(M)
    assert len({g.tobytes() for g in G}) == 8 * M * M
    ident = q.symmetry.neighbours(q.symmetry.plot(q.symmetry.translations(M)[0], M), M)
    assert all(q.symmetry.neighbours(q.symmetry.plot(g, M), M) == ident for g in G)


@pytest.mark.parametrize("kind,L,kw", [("crbm", 6, dict(k=5, alpha=4)), ("dcrbm", 8, dict(k=3, layers=[4, 6, 4]))])
def test_sym_logpsi_energy_gradient(kind, L, kw):
    q, sm, om = _pair(kind, L, 2e-1, 41, **kw)
    shape = (L, L)
    rng = np.random.default_rng(1)
    s = (rng.integers(0, 2, (5, L * L)) * 2 - 1).astype(np.int32)
    st = torch.as_tensor(s, device="cuda")
    want = osym.log_mean_exp(osym.log_psi_images(om, s, shape))
    got = sm.log_psi(st, shape).cpu().numpy()
    assert np.abs(np.exp(got - want) - 1).max() < 2e-5          # amplitudes agree (phase mod 2 pi)
    e_t = q.ising_energy(sm, st, system_shape=shape, H=0.8).cpu().numpy()
    assert np.abs(e_t - osym.sym_local_energy(om, s, shape, "tfim", H=0.8)).max() < 2e-5 * max(1, np.abs(e_t).max())
    if om.r + 1 <= L:
        e_h = q.heisenberg_energy(sm, st, system_shape=shape).cpu().numpy()
        assert np.abs(e_h - osym.sym_local_energy(om, s, shape, "heis")).max() < 2e-5 * max(1, np.abs(e_h).max())
    # gradient of sum_n Re[w_n conj(log psi_sym,n)] against finite differences of the oracle
    w = (rng.standard_normal(5) + 1j * rng.standard_normal(5)) / 5
    g = q.logpsi_gradient(sm, st, torch.as_tensor(w.astype(np.complex64), device="cuda"), shape).cpu().numpy()
    flat = om.flat_params().copy()
    f = lambda: float(np.real((w * np.conj(osym.log_mean_exp(osym.log_psi_images(om, s, shape)))).sum()))
    for i in rng.choice(flat.size, 10, replace=False):
        d = np.zeros_like(flat); d[i] = 1e-5
        om.set_flat_params(flat + d); fp = f()
        om.set_flat_params(flat - d); fm = f()
        assert abs((fp - fm) / 2e-5 - g[i]) < 2e-4 * max(1.0, np.abs(g).max()), i
    om.set_flat_params(flat)


@pytest.mark.parametrize("num_flips", [1, 2])
@pytest.mark.parametrize("L,k,alpha,S,n_steps", [(6, 3, 3, 20, 120), (10, 5, 4, 12, 80)], ids=["6x6_k3a3", "C4_10x10_k5a4"])
def test_sym_sweep_lockstep(num_flips, L, k, alpha, S, n_steps):
    """Symmetric Metropolis sweep vs the oracle's brute-force psi_sym sampler, in lock-step with the float64 oracle: a
    small case and BASELINE config C4's own shape (10x10, CRBM k = 5, alpha = 4, both flip counts)."""
    q, sm, om = _pair("crbm", L, 3e-1, 43, k=k, alpha=alpha)
    shape = (L, L)
    rng = np.random.default_rng(2)
    init = (rng.integers(0, 2, (S, L * L)) * 2 - 1).astype(np.int32)
    pos = rng.integers(0, L * L, (n_steps, S, num_flips)).astype(np.int32)
    u = rng.random((n_steps, S)).astype(np.float32)
    if num_flips == 2:
        pos[2, 1] = pos[2, 1, 0]
    smp = q.Sampler(sm, shape, sm.r, S, num_flips)
    smp.feed(init, pos, u)
    smp.mcmc_op(n_its=n_steps, trace=True)
    acc = smp.accept_trace.cpu().numpy().astype(bool)
    lr = smp.logratio_trace.cpu().numpy()
    osmp = osym.SymSampler(om, shape, num_flips)
    osmp.reset(init)
    from test_gpu_parity import tie_band, record_ties
    ties, err, worst, worst_ratio = 0, 0.0, 0.0, 0.0
    for i in range(n_steps):
        osmp.step(pos[i], u[i], force_mask=acc[i])
        t = osmp.last_log_ratio.real
        if num_flips == 2:
            t = np.where(pos[i, :, 0] == pos[i, :, 1], 0.0, t)
        err = max(err, float((np.abs(lr[i] - t) / np.maximum(1, np.abs(t))).max()))
        for c in np.nonzero(osmp.last_own_mask != acc[i])[0]:
            # a decision that differs from the float64 oracle's must be a tie inside the same band as the plain sweeps
            logu = np.log(max(float(u[i, c]), 1e-45))
            gap = abs(2 * t[c] - logu)
            assert gap < tie_band(2 * t[c], logu), (i, c, gap)
            worst, worst_ratio = max(worst, gap), max(worst_ratio, gap / tie_band(2 * t[c], logu))
            ties += 1
    record_ties("sym-%dx%d-k%d-a%d-flips%d" % (L, L, k, alpha, num_flips), ties, worst, worst_ratio, err, 0.0, n_steps * S)
    assert err < 2e-5 and ties <= 2
    assert np.array_equal(smp.spins.cpu().numpy().astype(np.int32), osmp.states)
    assert 0.05 < acc.mean() < 0.99
    if num_flips == 2:
        assert acc[2, 1]


@pytest.mark.parametrize("kind,L,kw", [("crbm", 10, dict(k=5, alpha=4)), ("dcrbm", 8, dict(k=3, layers=[4, 6, 4]))])
def test_sym_one_launch_paths_equal_image_composition(kind, L, kw):
    """north_star (3): the symmetry average is folded into the launches (image = blockIdx.y).  The entry points
    qmc_logpsi_forward_sym / qmc_local_energy_sym / qmc_logpsi_backward_sym must agree with composing 8 plain calls
    per image in torch (complex128 softmax), and stay within their launch budgets (2 / 3 / 4 + the parameter repack)."""
    q, sm, om = _pair(kind, L, 2e-1, 47, **kw)
    from qmcnn_b200 import _lib
    lib = _lib.load()
    shape = (L, L)
    rng = np.random.default_rng(5)
    N = 37
    st = torch.as_tensor((rng.integers(0, 2, (N, L * L)) * 2 - 1).astype(np.int8), device="cuda")
    w = torch.as_tensor(((rng.standard_normal(N) + 1j * rng.standard_normal(N)) / N).astype(np.complex64), device="cuda")

    def launches(fn):
        n0 = lib.qmc_launch_count()
        sm._bind(shape)                                  # the parameter repack launches are not part of the budget
        n_bind = lib.qmc_launch_count() - n0
        n0 = lib.qmc_launch_count()
        out = fn()
        return out, lib.qmc_launch_count() - n0 - n_bind

    lp, n = launches(lambda: sm.log_psi(st, shape))
    assert n <= 2, n
    want = sm.log_psi_composed(st, shape)
    # the composition carries each image's log psi as an fp32 total (|log psi| ~ 1e2: ulp ~ 8e-6 per image); the launch
    # sums per-site factors in double, and is the one held to the float64 oracle at 2e-5 in the test above
    assert (torch.exp(lp.to(torch.complex128) - want.to(torch.complex128)) - 1).abs().max().item() < 2e-4
    for ham, fn, kwargs in (("tfim", q.ising_energy, dict(H=0.7)), ("heis", q.heisenberg_energy, {})):
        if ham == "heis" and om.r + 1 > L:
            continue
        e, n = launches(lambda: fn(sm, st, system_shape=shape, **kwargs))
        assert n <= 3, (ham, n)
        want = sm.local_energy_composed(fn, st, shape, **kwargs)
        # element by element: where the images nearly cancel in sum_g psi_g, |E_loc| reaches 1e3-1e4 and both sides carry
        # the fp32 rounding of the image amplitudes amplified by that cancellation
        assert ((e - want).abs() / want.abs().clamp(min=1.0)).max().item() < 2e-4, ham
    g, n = launches(lambda: q.logpsi_gradient(sm, st, w, shape))
    assert n <= 4, n
    want = sm.gradient_composed(lambda im, s_, w_, sh: q.logpsi_gradient(im, s_, w_, sh), st, w, shape)
    assert (g - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())
    # moments accumulate in the same call
    mom = torch.zeros(4, dtype=torch.float64, device="cuda")
    e = q.ising_energy(sm, st, system_shape=shape, H=0.7, moments=mom)
    assert mom[0].item() == N and abs(mom[1].item() - e.real.double().sum().item()) < 1e-6 * N
