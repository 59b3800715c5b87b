"""The oracle against golden vectors recorded from the reference's own Python
(tests/golden/*.npz, produced by tests/golden/make_golden.py: /root/reference's
helpers.py / models.py / sampler.py / mcmc_tf.py executed unmodified over the eager
TF-1 API shim in oracle/tf1_shim).  CPU only; nothing here reads /root/reference.

Bars: integers (samples, accept decisions, window indices, bookkeeping) bit-exact;
float64 oracle vs the float64 reference run 1e-10; float32 oracle vs the float32
reference run 1e-5 relative (north_star's tolerance).
"""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "c1_tfim_crbm": dict(model=("CRBM", 5, 4, 2), shape=(6, 6), ham="tfim", H=1.0, num_samples=16, num_flips=1),
    "heis_crbm": dict(model=("CRBM", 5, 4, 2), shape=(6, 6), ham="heisenberg", H=1.0, num_samples=16, num_flips=2),
    "tfim_dcrbm": dict(model=("DCRBM", 3, (4, 4, 2), 2), shape=(8, 8), ham="tfim", H=3.0, num_samples=8, num_flips=1),
    "tfim_dcrbm888": dict(model=("DCRBM", 3, (8, 8, 8), 2), shape=(8, 8), ham="tfim", H=1.0, num_samples=8, num_flips=1),
    "tfim_crbm_1d": dict(model=("CRBM", 3, 2, 1), shape=(10,), ham="tfim", H=1.0, num_samples=8, num_flips=1),
    "heis_dcrbm_1d": dict(model=("DCRBM", 3, (4, 2), 1), shape=(9,), ham="heisenberg", H=1.0, num_samples=8, num_flips=2),
    "tfim_crbm_3d": dict(model=("CRBM", 3, 2, 3), shape=(5, 5, 5), ham="tfim", H=1.0, num_samples=4, num_flips=1),
    "tfim_crbm_sps2": dict(model=("CRBM", 5, 4, 2), shape=(6, 6), ham="tfim", H=1.0, num_samples=8, num_flips=1,
                           max_num_samplers=4),
    "heis_dcrbm": dict(model=("DCRBM", 3, (4, 2), 2), shape=(6, 6), ham="heisenberg", H=1.0, num_samples=8,
                       num_flips=2),
}


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def build_model(spec, gold, dtype, prefix="param/"):
    if spec[0] == "CRBM":
        m = oracle.CRBM(spec[1], (spec[1] - 1) // 2, spec[2], spec[3], dtype=dtype)
    else:
        m = oracle.DCRBM(spec[1], list(spec[2]), spec[3], dtype=dtype)
    for n in m.names:
        assert gold[prefix + n].shape == m.params[n].shape, n      # HWIO shapes, creation order
        m.params[n] = gold[prefix + n].astype(dtype)
    return m


def energy(case, model, states):
    if case["ham"] == "tfim":
        return oracle.ising_energy(model, states, case["shape"], model.r, H=case["H"])
    return oracle.heisenberg_energy(model, states, case["shape"], model.r)


def run_sampler(case, model, init, pos, u, new_samples=True, smp=None):
    if smp is None:
        cls = type("S", (oracle.Sampler,), {"MAX_NUM_SAMPLERS": case.get("max_num_samplers", 1000)})
        smp = cls(model, case["shape"], model.r, case["num_samples"], case["num_flips"])
    smp.new_samples = new_samples
    S = smp.num_samplers
    smp.mcmc_reset(init.reshape((S,) + tuple(case["shape"])).astype(np.int32), pos.astype(np.int32), u)
    acc = np.zeros((smp.sample_its, S), np.uint8)
    for i in range(smp.sample_its):
        smp.mcmc_step(i)
        acc[i] = smp.last_mask
    return smp, acc


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sampler_reproduces_reference_chain(name, dtype):
    """Sampler.mcmc_op (sampler.py:72-177): bookkeeping, reset factors, every accept
    decision, final padded state and the sample matrix in the reference's row order."""
    case, g = CASES[name], load(name)
    model = build_model(case["model"], g, dtype)
    smp, acc = run_sampler(case, model, g["initial_states"], g["flip_positions"], g["accept_sample"])
    bk = [smp.num_samplers, smp.its_per_sample, smp.samples_per_sampler, smp.therm_its, smp.sample_its,
          smp.padded_size]
    assert bk == list(g["bookkeeping"])
    assert np.array_equal(acc, g["accept"])
    assert np.array_equal(smp.current_samples, g["final_current_samples"])
    assert np.array_equal(smp.samples.reshape(case["num_samples"], -1), g["samples"])
    key, tol = ("final_factors", 1e-10) if dtype == np.float64 else ("f32/final_factors", 2e-5)
    assert np.abs(smp.current_factors - g[key]).max() <= tol * np.abs(g[key]).max()


@pytest.mark.parametrize("name", sorted(CASES))
def test_factors_and_energies(name):
    """model.factors (models.py:31-67 / 95-131) and ising_energy / heisenberg_energy
    (mcmc_tf.py:59-141) on the reference's samples."""
    case, g = CASES[name], load(name)
    shape, r = case["shape"], None
    for dtype, pre, tol in ((np.float64, "", 1e-10), (np.float32, "f32/", 1e-5)):
        model = build_model(case["model"], g, dtype)
        r = model.r
        states = g["samples"].astype(np.int32)
        padded = oracle.pad(states.reshape((-1,) + tuple(shape)), shape, [(r - 1) // 2] * len(shape))
        assert np.array_equal(padded[0], g["padded_samples_row0"])
        f = model.factors(padded)
        assert f.shape == g[pre + "factors"].shape
        assert np.abs(f - g[pre + "factors"]).max() <= tol * np.abs(g[pre + "factors"]).max()
        e = energy(case, model, states)
        assert np.abs(e - g[pre + "energies"]).max() <= tol * np.abs(g[pre + "energies"]).max()
        loss = oracle.loss_op(f, e)
        # loss_op is a covariance: a difference of products of size |E| |log psi|, which sets the error scale
        scale = np.abs(e).max() * np.abs(f.reshape(f.shape[0], -1).sum(1)).max()
        assert abs(loss - g[pre + "loss"]) <= tol * max(1.0, scale)


@pytest.mark.parametrize("name", sorted(CASES))
def test_two_optimisation_iterations(name):
    """optimize_op twice (mcmc_tf.py:156-179, 216-221): fresh chains, then persistent chains
    under the updated parameters; gradient of loss_op; TF-1 Adam."""
    case, g = CASES[name], load(name)
    model = build_model(case["model"], g, np.float64)
    names = model.names
    smp = None
    m = np.zeros_like(model.flat_params())
    v = np.zeros_like(m)
    for it in range(2):
        pre = "opt%d/" % it
        smp, _ = run_sampler(case, model, g[pre + "initial_states"], g[pre + "flip_positions"],
                             g[pre + "accept_sample"], new_samples=(it == 0), smp=smp)
        samples = smp.samples.reshape(case["num_samples"], -1)
        assert np.array_equal(samples, g[pre + "samples"]), "iteration %d samples" % it
        e = energy(case, model, samples)
        assert np.abs(e - g[pre + "energies"]).max() <= 1e-10 * np.abs(g[pre + "energies"]).max()
        r = model.r
        padded = oracle.pad(samples.reshape((-1,) + tuple(case["shape"])), case["shape"],
                            [(r - 1) // 2] * len(case["shape"]))
        grad, _ = oracle.vmc_gradient(model, padded, e)
        want = np.concatenate([g[pre + "grad/" + n].ravel() for n in names])
        assert np.abs(grad - want).max() <= 1e-9 * np.abs(want).max()
        p, m, v = oracle.adam_tf1_step(model.flat_params(), grad, m, v, it + 1, lr=3e-3)
        want_p = np.concatenate([g[pre + "param/" + n].ravel() for n in names])
        assert np.abs(p - want_p).max() <= 1e-12
        model.set_flat_params(p)


def test_helpers_match_reference():
    """helpers.py: create_index_matrix, all_windows, interactions, pad/unpad, gather_windows,
    update_windows in 1-D, 2-D, 3-D, odd and even windows, windows larger than the lattice."""
    g = np.load(os.path.join(GOLD, "helpers.npz"))
    specs = {"1d": ((7,), (3,)), "2d": ((4, 5), (3, 3)), "2d_even": ((6, 6), (4, 4)),
             "3d": ((3, 4, 3), (3, 3, 3)), "2d_wrap": ((3, 3), (5, 5))}
    for tag, (shape, win) in specs.items():
        x, s = g[tag + "/x"], g[tag + "/s"]
        assert np.array_equal(oracle.create_index_matrix(shape, win), g[tag + "/index_matrix"]), tag
        assert np.array_equal(oracle.all_windows(x, shape, win), g[tag + "/all_windows"]), tag
        assert np.array_equal(oracle.interactions(s, shape), g[tag + "/interactions"]), tag
        p = tuple(g[tag + "/pad_size"])
        padded = oracle.pad(x.reshape((3,) + shape), shape, p)
        assert np.array_equal(padded, g[tag + "/padded"]), tag
        assert np.array_equal(oracle.unpad(padded, p), g[tag + "/unpadded"]), tag
        assert np.array_equal(oracle.gather_windows(x, g[tag + "/centers"], shape, win),
                              g[tag + "/gather_windows"]), tag
        if tag + "/update_windows" in g:
            got = oracle.update_windows(x, g[tag + "/centers"], g[tag + "/updates"], g[tag + "/mask"], shape, win)
            assert np.array_equal(got, g[tag + "/update_windows"]), tag


def test_product_helper_shims_match_reference():
    """The PRODUCT's helper API (qmcnn_b200.helpers: create_index_matrix, all_windows, interactions, pad / unpad,
    gather_windows, update_windows - eager torch, any device) against the same reference-run vectors."""
    import torch
    import qmcnn_b200.helpers as ph
    g = np.load(os.path.join(GOLD, "helpers.npz"))
    specs = {"1d": ((7,), (3,)), "2d": ((4, 5), (3, 3)), "2d_even": ((6, 6), (4, 4)),
             "3d": ((3, 4, 3), (3, 3, 3)), "2d_wrap": ((3, 3), (5, 5))}
    t = lambda a: torch.as_tensor(np.asarray(a))
    for tag, (shape, win) in specs.items():
        x, s = g[tag + "/x"], g[tag + "/s"]
        im = ph.create_index_matrix(shape, win)
        assert im.dtype == np.int32 and np.array_equal(im, g[tag + "/index_matrix"]), tag
        assert np.array_equal(ph.all_windows(t(x), shape, win).numpy(), g[tag + "/all_windows"]), tag
        assert np.array_equal(ph.interactions(t(s), shape).numpy(), g[tag + "/interactions"]), tag
        p = tuple(int(v) for v in g[tag + "/pad_size"])
        padded = ph.pad(t(x).reshape((3,) + shape), shape, p)
        assert np.array_equal(padded.numpy(), g[tag + "/padded"]), tag
        assert np.array_equal(ph.unpad(padded, p).numpy(), g[tag + "/unpadded"]), tag
        if tag + "/centers" in g:
            assert np.array_equal(ph.gather_windows(t(x), t(g[tag + "/centers"]), shape, win).numpy(),
                                  g[tag + "/gather_windows"]), tag
        if tag + "/update_windows" in g:
            got = ph.update_windows(t(x).clone(), t(g[tag + "/centers"]), t(g[tag + "/updates"]),
                                    t(g[tag + "/mask"]), shape, win)
            assert np.array_equal(got.numpy(), g[tag + "/update_windows"]), tag


def test_factors_in_1d_2d_3d():
    """models.py:56-61 / 118-123: the conv1d / conv2d / conv3d branches."""
    g = np.load(os.path.join(GOLD, "factors_nd.npz"))
    specs = {"crbm1d": ("CRBM", 3, 2, 1), "crbm2d": ("CRBM", 5, 4, 2), "crbm3d": ("CRBM", 3, 2, 3),
             "dcrbm1d": ("DCRBM", 3, (4, 4, 2), 1), "dcrbm2d": ("DCRBM", 3, (4, 4, 2), 2),
             "dcrbm3d": ("DCRBM", 3, (4, 2), 3)}
    for tag, spec in specs.items():
        model = build_model(spec, g, np.float64, prefix=tag + "/param/")
        s = g[tag + "/spins"].astype(np.int32)
        shape = s.shape[1:]
        padded = oracle.pad(s, shape, [(model.r - 1) // 2] * len(shape))
        f = model.factors(padded)
        assert np.abs(f - g[tag + "/factors"]).max() <= 1e-10 * np.abs(g[tag + "/factors"]).max(), tag


def test_symmetry_group_matches_reference_notebook():
    """symmetry.ipynb cell 0 executed as it is (its assertions passed while recording): D4, T and G = D4 x T in
    the notebook's element order, the site permutation `plot(g)` of every group element, the identity's
    neighbour table, and cells 1-2 - reproduced by the oracle's and the product's group utilities."""
    import qmcnn_b200.symmetry as ps
    from oracle import symmetry as osym
    g = np.load(os.path.join(GOLD, "symmetry.npz"))
    M = int(g["M"])
    for mod_ in (osym, ps):
        assert np.array_equal(mod_.d4(M), g["D4"])
        assert np.array_equal(mod_.translations(M), g["T"])
        G = mod_.group(M)
        assert np.array_equal(G, g["G"]) and len(G) == 8 * M * M
        assert np.array_equal(np.stack([mod_.plot(x, M) for x in G]), g["plots"])
        idn = mod_.neighbours(mod_.plot(mod_.translations(M)[0], M), M)
        assert np.array_equal(np.array([idn[i] for i in range(M * M)]), g["identity_neighbours"])
        grid = mod_.plot(mod_.mod(np.dot(mod_.d4(M)[7], mod_.translations(M)[5]), M), M)
        assert np.array_equal(grid, g["cell1_grid"])
        assert list(mod_.neighbours(grid, M)[12]) == list(g["cell2_neighbours_12"])
