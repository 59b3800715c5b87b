"""The CUDA path against golden vectors recorded from the reference's own Python
(tests/golden/*.npz; see tests/golden/make_golden.py and oracle/tf1_shim): the
reference's Sampler.mcmc_op, model.factors, ising_energy / heisenberg_energy, the
gradient of loss_op and the TF-1 Adam update, on identical parameters, initial
lattices, proposals and uniforms.

Bars (north_star): accept decisions, final states and samples bit-exact (the recorded
chains are free of float32 near-ties by construction); log psi factors and local
energies 1e-5 relative; gradient 1e-4 of its scale.
"""
import os

import numpy as np
import pytest
import torch

import qmcnn_b200 as q
from test_golden_oracle import CASES, GOLD, load

pytestmark = pytest.mark.gpu


def build(spec, gold, prefix="param/"):
    if spec[0] == "CRBM":
        m = q.CRBM(spec[1], (spec[1] - 1) // 2, spec[2], spec[3], seed=0)
    else:
        m = q.DCRBM(spec[1], list(spec[2]), spec[3], seed=0)
    set_params(m, gold, prefix)
    return m


def set_params(m, gold, prefix):
    flat = np.concatenate([np.asarray(gold[prefix + n], np.float32).ravel() for n in m.names])
    m.set_flat_params(flat)
    for n in m.names:       # named HWIO views agree with the reference's variables
        assert np.array_equal(m.params[n].cpu().numpy(), np.asarray(gold[prefix + n], np.float32)), n


def make_sampler(case, model):
    cls = type("S", (q.Sampler,), {"MAX_NUM_SAMPLERS": case.get("max_num_samplers", 1000)})
    return cls(model, case["shape"], model.r, case["num_samples"], case["num_flips"])


def energy(case, model, states):
    if case["ham"] == "tfim":
        return q.ising_energy(model, states, system_shape=case["shape"], H=case["H"])
    return q.heisenberg_energy(model, states, system_shape=case["shape"])


SWEEPABLE = sorted(CASES)       # incl. heis_dcrbm: two flips of a deep model run the generic full-forward path


@pytest.fixture(params=["default", "inplace", "classic", "inplace-rowmajor"])
def sweep_path(request):
    """Every sweep decomposition must reproduce the reference's chains: the default choice, the in-place
    persistent kernel forced on (k_sweep_ip; it is the default only for big models; with the conflict-free site
    deal and with the row-major order) and the classic ping-pong kernel (each falls back to the classic
    kernel for shapes it does not cover, e.g. CRBM or two flips).  Returns the model tuning dict."""
    return {"default": {}, "inplace": dict(flags=q.FLAG_SWEEP_INPLACE), "classic": dict(flags=q.FLAG_SWEEP_CLASSIC),
            "inplace-rowmajor": dict(flags=q.FLAG_SWEEP_INPLACE | q.FLAG_IP_ROWMAJOR_SITES)}[request.param]


@pytest.mark.parametrize("name", SWEEPABLE)
def test_mcmc_op_reproduces_reference_chain(name, sweep_path):
    case, g = CASES[name], load(name)
    model = build(case["model"], g)
    model.tuning = dict(sweep_path)
    smp = make_sampler(case, model)
    assert [smp.num_samplers, smp.its_per_sample, smp.samples_per_sampler, smp.therm_its, smp.sample_its,
            smp.padded_size] == list(g["bookkeeping"])
    smp.feed(g["initial_states"], g["flip_positions"].astype(np.int32), g["accept_sample"])
    samples = smp.mcmc_op(trace=True)
    acc = smp.accept_trace.cpu().numpy()
    bad = np.argwhere(acc != g["accept"])
    assert bad.size == 0, "first differing decision (step, chain) = %s" % (bad[:1],)
    assert samples.dtype == torch.int32 and tuple(samples.shape) == g["samples"].shape
    assert np.array_equal(samples.cpu().numpy(), g["samples"])
    assert np.array_equal(smp.current_samples_var.cpu().numpy(), g["final_current_samples"])
    f = smp.current_factors_var.cpu().numpy()
    assert np.abs(f - g["final_factors"]).max() <= 1e-5 * np.abs(g["final_factors"]).max()
    # log-ratios of the accepted moves against the reference's own factor differences
    lr = smp.logratio_trace.cpu().numpy()
    ok = ~np.isnan(g["logratio_re"])
    assert np.abs(lr[ok] - g["logratio_re"][ok]).max() <= 2e-5 * max(1.0, np.abs(g["logratio_re"][ok]).max())


def test_two_flip_deep_model_uses_the_generic_path():
    """sampler.py:106-122 samples any model with num_flips = 2.  Two independent uniform sites of a deep model
    have a bounding box + receptive field wider than the lattice, which the incremental kernels do not cover:
    the Sampler then runs the generic full-forward path (the reference's own algorithm) instead of raising."""
    case, g = CASES["heis_dcrbm"], load("heis_dcrbm")
    model = build(case["model"], g)
    smp = make_sampler(case, model)
    assert smp._nd, "expected the generic path for a two-flip DCRBM on 6x6"


@pytest.mark.parametrize("name", sorted(CASES))
def test_factors_and_energies(name):
    case, g = CASES[name], load(name)
    model = build(case["model"], g)
    shape, halo = case["shape"], (model.r - 1) // 2
    states = torch.as_tensor(g["samples"].astype(np.int32), device="cuda")
    padded = q.pad(states.reshape((-1,) + tuple(shape)), shape, [halo] * len(shape))
    assert np.array_equal(padded[0].cpu().numpy(), g["padded_samples_row0"])
    f = model.factors(padded).cpu().numpy()
    assert f.shape == g["factors"].shape
    assert np.abs(f - g["factors"]).max() <= 1e-5 * np.abs(g["factors"]).max()
    e = energy(case, model, states).cpu().numpy()
    assert np.abs(e - g["energies"]).max() <= 1e-5 * np.abs(g["energies"]).max()
    eb = q.batched_op(lambda s: energy(case, model, s), states, case["num_samples"] // 2).cpu().numpy()
    assert np.array_equal(eb, e)
    loss = float(q.loss_op(torch.as_tensor(f, device="cuda"), torch.as_tensor(e, device="cuda")))
    scale = np.abs(e).max() * np.abs(f.reshape(f.shape[0], -1).sum(1)).max()
    assert abs(loss - float(g["loss"])) <= 1e-5 * max(1.0, scale)


@pytest.mark.parametrize("name", SWEEPABLE)
def test_two_optimisation_iterations(name):
    """optimize_op(...).run() twice: fresh chains, then persistent chains under the updated
    parameters (mcmc_tf.py:156-179, 216-221).  Each iteration starts from the reference's
    float32 parameters so that the comparison does not compound."""
    case, g = CASES[name], load(name)
    model = build(case["model"], g)
    smp = make_sampler(case, model)
    step = q.optimize_op(smp, model, lambda s: energy(case, model, s), learning_rate=3e-3)
    for it in range(2):
        pre = "opt%d/" % it
        if it:
            set_params(model, g, "f32/opt%d/param/" % (it - 1))
        before = model.flat.clone()
        smp.flip_positions_var = smp.accept_sample_var = None
        smp.feed(g[pre + "initial_states"], g[pre + "flip_positions"].astype(np.int32), g[pre + "accept_sample"])
        e = step.run(new_samples=(it == 0)).cpu().numpy()
        assert np.array_equal(smp.samples_int8().cpu().numpy(), g[pre + "samples"]), "iteration %d" % it
        assert np.abs(e - g[pre + "energies"]).max() <= 1e-5 * np.abs(g[pre + "energies"]).max()
        want = np.concatenate([g[pre + "grad/" + n].ravel() for n in model.names])
        grad = step.last_grad.cpu().numpy()
        assert np.abs(grad - want).max() <= 1e-4 * np.abs(want).max()
        # Adam arithmetic on the reference's own gradient (m/sqrt(v) amplifies gradient noise
        # near zero crossings, so the update rule is checked with identical inputs)
        opt = q.AdamTF1(before.clone(), 3e-3)
        opt.t = it
        if it:
            g0 = np.concatenate([g["f32/opt0/grad/" + n].ravel() for n in model.names]).astype(np.float32)
            g0 = torch.as_tensor(g0, device="cuda")
            opt.m = 0.1 * g0
            opt.v = 0.001 * g0 * g0
        gref = np.concatenate([g["f32/" + pre + "grad/" + n].ravel() for n in model.names]).astype(np.float32)
        opt.step(torch.as_tensor(gref, device="cuda"))
        want_p = np.concatenate([g["f32/" + pre + "param/" + n].ravel() for n in model.names])
        assert np.abs(opt.flat.cpu().numpy() - want_p).max() <= 2e-6


def test_eval_op_and_optimize_op_return_shapes():
    """mcmc_tf.py:182-194 eval_op = batched energies of a fresh mcmc_op; mcmc_tf.py:157-179 optimize_op returns
    (energies, train_op) - the product's object unpacks into the two and `run` is the eager sess.run."""
    name = "c1_tfim_crbm"
    case, g = CASES[name], load(name)
    model = build(case["model"], g)
    smp = make_sampler(case, model)
    smp.feed(g["initial_states"], g["flip_positions"].astype(np.int32), g["accept_sample"])
    fn = lambda s: energy(case, model, s)
    e = q.eval_op(smp, model, fn, batch_size=case["num_samples"] // 2).cpu().numpy()
    assert np.array_equal(smp.samples_int8().cpu().numpy(), g["samples"])
    assert np.abs(e - g["energies"]).max() <= 1e-5 * np.abs(g["energies"]).max()
    opt = q.optimize_op(smp, model, fn)
    energies, train_op = opt
    assert len(opt) == 2 and train_op is opt and opt[0] is energies
    with pytest.raises(q.QmcError):
        energies.eval()
    e2, _ = q.mcmc.run(opt, feed_dict={"new_samples": True})
    assert torch.equal(energies.eval(), e2) and e2.shape == (case["num_samples"],)
