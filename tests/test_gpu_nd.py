"""1-D and 3-D lattices (models.py:56-61 / 118-123 conv1d / conv3d branches; n_dims-generic sampler.py and
mcmc_tf.py) through the generic CUDA path (qmc_nd_*): model.factors against the reference-run golden vectors,
the sampler in lock-step with the oracle, both energy estimators against the oracle; n_dims = 2 factors
through the same path are cross-checked against the golden vectors too."""
import os

import numpy as np
import pytest
import torch

import oracle
import qmcnn_b200 as q

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TIE_BAND = 2e-5


def pair(kind, n_dims, scale, seed, k=3, alpha=2, layers=(4, 4, 2), dtype=np.float32):
    rng = np.random.default_rng(seed)
    if kind == "crbm":
        om = oracle.CRBM(k, (k - 1) // 2, alpha, n_dims, rng=rng, scale=scale, dtype=dtype)
        gm = q.CRBM(k, (k - 1) // 2, alpha, n_dims, seed=0)
    else:
        om = oracle.DCRBM(k, list(layers), n_dims, rng=rng, scale=scale, dtype=dtype)
        gm = q.DCRBM(k, list(layers), n_dims, seed=0)
    gm.set_flat_params(om.flat_params().astype(np.float32))
    return gm, om


SPECS = {"crbm1d": ("CRBM", 3, 2, 1), "crbm2d": ("CRBM", 5, 4, 2), "crbm3d": ("CRBM", 3, 2, 3),
         "dcrbm1d": ("DCRBM", 3, (4, 4, 2), 1), "dcrbm2d": ("DCRBM", 3, (4, 4, 2), 2), "dcrbm3d": ("DCRBM", 3, (4, 2), 3)}


@pytest.mark.parametrize("tag", sorted(SPECS))
def test_factors_match_reference_run_vectors(tag):
    """model.factors in 1-D, 2-D and 3-D against what the reference's own models.py computed (float64)."""
    g = np.load(os.path.join(GOLD, "factors_nd.npz"))
    spec = SPECS[tag]
    m = (q.CRBM(spec[1], (spec[1] - 1) // 2, spec[2], spec[3], seed=0) if spec[0] == "CRBM"
         else q.DCRBM(spec[1], list(spec[2]), spec[3], seed=0))
    m.set_flat_params(np.concatenate([g["%s/param/%s" % (tag, n)].ravel() for n in m.names]).astype(np.float32))
    for n in m.names:
        assert tuple(m.params[n].shape) == g["%s/param/%s" % (tag, n)].shape, n      # [k]*n_dims + [C_in, C_out]
    s = torch.as_tensor(g[tag + "/spins"].astype(np.int32), device="cuda")
    shape = tuple(s.shape[1:])
    halo = (m.r - 1) // 2
    x = q.pad(s, shape, [halo] * len(shape))
    want = g[tag + "/factors"]
    if m.n_dims == 2:          # the generic path on a 2-D lattice, and the tuned kernels
        f_nd = m.nd_forward(s.reshape(s.shape[0], -1), shape)[0].view((-1,) + shape).cpu().numpy()
        assert np.abs(f_nd - want).max() <= 1e-5 * np.abs(want).max()
    f = m.factors(x).cpu().numpy()                 # general VALID-conv path (no assumption about x)
    assert f.shape == want.shape
    assert np.abs(f - want).max() <= 1e-5 * np.abs(want).max()
    fp = m.factors(x, periodic=True).cpu().numpy()  # caller-guaranteed wrap-padded image: inner lattice only
    assert np.abs(fp - want).max() <= 1e-5 * np.abs(want).max()
    lp = m.log_psi(s.reshape(s.shape[0], -1), shape).cpu().numpy()
    assert np.abs(lp - want.reshape(want.shape[0], -1).sum(1)).max() <= 1e-5 * np.abs(want).sum(tuple(range(1, want.ndim))).max()
    if halo:
        # not a periodic image: models.py:31-67 is a plain VALID conv and must still answer; flipping a halo
        # corner changes exactly the factors whose receptive field contains it
        bad = x.clone()
        bad[(0,) + (0,) * len(shape)] *= -1
        fb = m.factors(bad).cpu().numpy()
        assert fb.shape == want.shape
        changed = np.abs(fb - f).reshape(f.shape[0], -1).max(1)
        assert changed[0] > 0 and np.all(changed[1:] == 0)


@pytest.mark.parametrize("kind,shape,flips", [("crbm", (12,), 1), ("dcrbm", (11,), 1), ("crbm", (4, 3, 4), 1),
                                              ("dcrbm", (5, 5, 5), 1), ("crbm", (10,), 2), ("crbm", (3, 4, 3), 2)])
def test_sweep_lockstep_with_oracle(kind, shape, flips):
    """Sampler.mcmc_op on fed-in randoms: every accept decision equals the float64 oracle's except inside the
    float32 tie band; final states equal; sample write-out order as in sampler.py:135-152."""
    n_dims, n = len(shape), int(np.prod(shape))
    layers = (4, 4, 2) if n_dims == 1 else (4, 2)
    gm, om = pair(kind, n_dims, 0.15 if kind == "crbm" else (0.7 if n_dims == 1 else 0.15), 5, layers=layers)
    om64 = om.astype(np.float64)
    S = 9
    GS = type("GS", (q.Sampler,), dict(MAX_NUM_SAMPLERS=S, SWEEPFACTOR=1, THERMFACTOR=1))
    OS = type("OS", (oracle.Sampler,), dict(MAX_NUM_SAMPLERS=S, SWEEPFACTOR=1, THERMFACTOR=1))
    gs = GS(gm, shape, om.r, 2 * S, flips)
    os64 = OS(om64, shape, om.r, 2 * S, flips)
    n_steps = gs.sample_its
    assert n_steps == os64.sample_its == 3 * n + 1
    rng = np.random.default_rng(17)
    init = (rng.integers(0, 2, (S,) + tuple(shape)) * 2 - 1).astype(np.int32)
    pos = rng.integers(0, n, (n_steps, S, flips)).astype(np.int32)
    if flips == 2:
        pos[1, 0] = pos[1, 0, 0]
    u = rng.random((n_steps, S)).astype(np.float32)
    gs.feed(init, pos, u)
    samples = gs.mcmc_op(trace=True)
    acc = gs.accept_trace.cpu().numpy().astype(bool)
    lr = gs.logratio_trace.cpu().numpy()
    os64.mcmc_reset(init, pos, u)
    ties = 0
    for i in range(n_steps):
        os64.mcmc_step(i, force_mask=acc[i])
        t = os64.last_log_ratio.real
        scale = np.maximum(1.0, np.abs(t))
        ident = (pos[i, :, 0] == pos[i, :, 1]) if flips == 2 else np.zeros(S, bool)
        assert (np.abs(lr[i] - t) / scale)[~ident].max() <= 2e-5
        for c in np.nonzero(os64.last_own_mask != acc[i])[0]:
            gap = abs(2.0 * float(t[c]) - np.log(max(float(u[i, c]), 1e-45))) / float(scale[c])
            assert gap < TIE_BAND, "step %d chain %d: decisions differ outside the tie band (%g)" % (i, c, gap)
            ties += 1
    assert ties <= 2 and 0 < acc.sum() < acc.size
    assert np.array_equal(gs.spins.cpu().numpy().astype(np.int32), os64.unpadded_current())
    assert np.array_equal(samples.cpu().numpy(), os64.samples.reshape(2 * S, n))
    assert np.array_equal(gs.current_samples_var.cpu().numpy(), os64.current_samples)
    f = gs.current_factors_var.cpu().numpy()
    assert np.abs(f - os64.current_factors).max() <= 2e-5 * max(1.0, np.abs(os64.current_factors).max())


@pytest.mark.parametrize("kind,shape", [("crbm", (12,)), ("dcrbm", (11,)), ("crbm", (5, 6, 5)), ("dcrbm", (7, 7, 7))])
def test_energies_match_oracle(kind, shape):
    """Lattice sides are >= K + 2 (K = receptive field): below that the reference's window trick sees an unflipped
    periodic image of a flipped site inside its (2K+1)-wide window (mcmc_tf.py:105-126) and no longer equals the
    flipped lattice, which is what this path evaluates."""
    n_dims, n = len(shape), int(np.prod(shape))
    layers = (4, 4, 2) if n_dims == 1 else (4, 2)
    gm, om = pair(kind, n_dims, 0.15 if kind == "crbm" else (0.7 if n_dims == 1 else 0.15), 9, layers=layers)
    om64 = om.astype(np.float64)
    states = (np.random.default_rng(3).integers(0, 2, (6 if n < 200 else 2, n)) * 2 - 1).astype(np.int32)
    st = torch.as_tensor(states, device="cuda")
    # 343 sites: every log-ratio is a float32 sum of 343 per-site differences of full-network factors (the
    # reference's own formulation); measured 1.2e-5, the float32 oracle is no better
    tol = 1e-5 if n < 200 else 3e-5
    e = q.ising_energy(gm, st, system_shape=shape, H=0.8).cpu().numpy()
    want = oracle.ising_energy(om64, states, shape, om.r, H=0.8)
    assert np.abs(e - want).max() <= tol * np.abs(want).max()
    e = q.heisenberg_energy(gm, st, system_shape=shape).cpu().numpy()
    want = oracle.heisenberg_energy(om64, states, shape, om.r)
    assert np.abs(e - want).max() <= tol * np.abs(want).max()
    eb = q.batched_op(lambda s: q.heisenberg_energy(gm, s, system_shape=shape), st, st.shape[0] // 2).cpu().numpy()
    assert np.array_equal(eb, e)


@pytest.mark.parametrize("kind,shape", [("crbm", (12,)), ("dcrbm", (11,)), ("crbm", (4, 3, 4)), ("dcrbm", (5, 5, 5)),
                                        ("dcrbm", (6, 5))])
def test_gradient_matches_autograd_oracle(kind, shape):
    """d loss_op / d params (mcmc_tf.py:35-56, 172-177) by the generic backward kernel against torch autograd of the
    oracle's restatement; n_dims = 2 through the same kernel against the tuned 2-D kernels as well."""
    n_dims, n = len(shape), int(np.prod(shape))
    layers = (4, 4, 2) if n_dims == 1 else (4, 2)
    gm, om = pair(kind, n_dims, 0.2, 21, layers=layers)
    rng = np.random.default_rng(2)
    N = 11
    states = (rng.integers(0, 2, (N, n)) * 2 - 1).astype(np.int32)
    e = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    w = torch.as_tensor(((e - e.mean()) / N).astype(np.complex64), device="cuda")
    st = torch.as_tensor(states, device="cuda")
    if n_dims == 2:
        import qmcnn_b200._lib as L
        d = gm.nd_desc(shape)
        lib = q.load_library()
        g = torch.zeros(gm.num_params, dtype=torch.float32, device="cuda")
        ws = torch.empty(lib.qmc_nd_backward_scratch_floats(d, 0, N), dtype=torch.float32, device="cuda")
        s8 = st.to(torch.int8).contiguous()
        L.check_nd(lib.qmc_nd_logpsi_backward(d, 0, gm.flat.data_ptr(), s8.data_ptr(), w.data_ptr(), N, ws.data_ptr(),
                                              g.data_ptr(), torch.cuda.current_stream().cuda_stream), "nd_backward")
        g = g.cpu().numpy()
        g2 = q.logpsi_gradient(gm, st, w, system_shape=shape).cpu().numpy()
        assert np.abs(g - g2).max() <= 2e-6 * np.abs(g2).max()
    else:
        g = q.logpsi_gradient(gm, st, w, system_shape=shape).cpu().numpy()
    xp = oracle.pad(states.reshape((N,) + tuple(shape)), shape, [(om.r - 1) // 2] * n_dims)
    want, _ = oracle.vmc_gradient(om.astype(np.float64), xp, e)
    assert np.abs(g - want).max() <= 1e-4 * np.abs(want).max()


def test_vmc_loop_on_a_chain_converges_to_exact_ground_state():
    """run_vmc on a 1-D TFIM chain of 10 spins at h = 1 (CRBM k = 3): the whole loop - generic-path sampler,
    energies, gradient, TF-1 Adam - approaches the exact ground-state energy per spin from above."""
    import itertools
    from scipy.sparse import lil_matrix
    from scipy.sparse.linalg import eigsh
    n = 10
    s = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=np.int32)
    idx = {tuple(r): i for i, r in enumerate(s)}
    Hm = lil_matrix((2 ** n, 2 ** n))
    bonds = (s * np.roll(s, -1, 1)).sum(1)
    for i, r in enumerate(s):
        Hm[i, i] = -bonds[i]
        for k in range(n):
            t = r.copy(); t[k] = -t[k]
            Hm[i, idx[tuple(t)]] = -1.0
    e0 = eigsh(Hm.tocsr(), k=1, which="SA")[0][0] / n
    model = q.CRBM(3, 1, 4, 1, seed=3)
    hist = q.run_vmc(model, (n,), "tfim", 1.0, num_samples=500, num_eval_samples=2000, optimization_its=200,
                     eval_freq=50, learning_rate=1e-2, energy_batch_size=1000, seed=5, log=None)
    last = [r for r in hist if "eval_energy" in r][-1]
    assert last["eval_energy"] < hist[0]["energy"] - 0.1
    assert last["eval_energy"] >= e0 - 5 * last["eval_stderr"] - 1e-3
    assert last["eval_energy"] <= e0 + 0.03 * abs(e0), "E = %.5f vs exact %.5f" % (last["eval_energy"], e0)
