"""World-size-2 gloo tests (CPU) of the data-parallel host logic: chain partition, the
energy-moment all-reduce and the gradient all-reduce reproduce the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle.philox import sweep_randoms


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        # qmcnn_b200.distributed is device-agnostic torch; load it without the CUDA package __init__
        spec = importlib.util.spec_from_file_location("qmc_dist", os.path.join(root, "qmcnn_b200", "distributed.py"))
        D = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(D)
        sys.path.insert(0, root)
        L, S_total = 4, 10
        rng = np.random.default_rng(5)
        model = oracle.CRBM(3, 1, 2, 2, rng=rng, scale=0.3, dtype=np.float64)
        states = (rng.integers(0, 2, (S_total, L * L)) * 2 - 1).astype(np.int32)
        first, count = D.chain_partition(S_total, rank, world)
        mine = states[first:first + count]
        e_local = torch.as_tensor(oracle.ising_energy(model, mine, (L, L), 3, H=0.9))
        n, mean, var, stderr = D.allreduce_energy_moments(e_local)
        w = D.vmc_weights(e_local, mean, n)
        # per-rank gradient of sum_n Re[w_n conj(dlogpsi_n)]: autograd of sum Re(w conj logpsi)
        xp = oracle.pad(mine.reshape(-1, L, L), (L, L), [1, 1])
        g_local = _weighted_grad(model, xp, w.numpy())
        g = D.allreduce_gradient(torch.as_tensor(g_local))
        out[rank] = dict(first=first, count=count, n=float(n), mean=complex(mean), var=float(var),
                         grad=g.numpy().copy())
    finally:
        dist.destroy_process_group()


def _weighted_grad(model, xp, w):
    """sum_n Re[w_n conj(d logpsi_n / dp)] by finite differences of sum_n Re[w_n conj(logpsi_n)]."""
    flat = model.flat_params().copy()
    g = np.zeros_like(flat)
    f = lambda: float(np.real((w * np.conj(model.log_psi(xp))).sum()))
    for i in range(flat.size):
        d = np.zeros_like(flat); d[i] = 1e-6
        model.set_flat_params(flat + d); fp = f()
        model.set_flat_params(flat - d); fm = f()
        g[i] = (fp - fm) / 2e-6
    model.set_flat_params(flat)
    return g


@pytest.mark.timeout(300)
def test_two_rank_reductions_match_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # single-process reference
    L, S_total = 4, 10
    rng = np.random.default_rng(5)
    model = oracle.CRBM(3, 1, 2, 2, rng=rng, scale=0.3, dtype=np.float64)
    states = (rng.integers(0, 2, (S_total, L * L)) * 2 - 1).astype(np.int32)
    e = oracle.ising_energy(model, states, (L, L), 3, H=0.9)
    xp = oracle.pad(states.reshape(-1, L, L), (L, L), [1, 1])
    g_ref, _ = oracle.vmc_gradient(model, xp, e)
    assert out[0]["first"] == 0 and out[0]["count"] == 5 and out[1]["first"] == 5 and out[1]["count"] == 5
    for r in range(world):
        assert out[r]["n"] == S_total
        assert abs(out[r]["mean"] - e.mean()) < 1e-12
        assert abs(out[r]["var"] - e.real.var()) < 1e-12
        assert np.abs(out[r]["grad"] - g_ref).max() < 1e-6      # FD noise only
    assert np.array_equal(out[0]["grad"], out[1]["grad"])       # replicas stay bit-identical


def test_chain_partition_covers_everything():
    spec_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import importlib.util
    spec = importlib.util.spec_from_file_location("qmc_dist", os.path.join(spec_root, "qmcnn_b200", "distributed.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    for total, world in ((32768, 8), (10, 3), (7, 8), (4096, 1)):
        seen = []
        for r in range(world):
            f, c = D.chain_partition(total, r, world)
            seen += list(range(f, f + c))
        assert seen == list(range(total))


def test_philox_streams_do_not_depend_on_the_partition():
    """Rank r feeds chain_id0 = first chain of its block: the union of the per-rank streams
    equals the single-rank stream (weak/strong scaling changes nothing statistically)."""
    total, steps = 12, 9
    pos_all, u_all = sweep_randoms(77, np.arange(total), 0, steps, 1, 400)
    for world in (2, 3, 4):
        for r in range(world):
            base, rem = divmod(total, world)
            cnt = base + (1 if r < rem else 0)
            first = r * base + min(r, rem)
            p, u = sweep_randoms(77, first + np.arange(cnt), 0, steps, 1, 400)
            assert np.array_equal(p, pos_all[:, first:first + cnt]) and np.array_equal(u, u_all[:, first:first + cnt])


def test_initial_lattices_do_not_depend_on_the_partition():
    """Fresh chains (sampler.py:74-79) are keyed by (seed, global chain id, reset count) like the proposals."""
    from oracle.philox import initial_spins
    total, n = 24, 400
    full = initial_spins(77, np.arange(total), n)
    assert full.shape == (total, n) and set(np.unique(full)) == {-1, 1}
    assert abs(full.mean()) < 0.05 and len({r.tobytes() for r in full}) == total
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            first, cnt = (r * (total // world) + min(r, total % world), total // world + (1 if r < total % world else 0))
            parts.append(initial_spins(77, first + np.arange(cnt), n))
        assert np.array_equal(np.concatenate(parts), full)
    assert not np.array_equal(initial_spins(77, np.arange(total), n, reset_index=1), full)
    assert not np.array_equal(initial_spins(78, np.arange(total), n), full)


class _RecordingSampler(object):
    MAX_NUM_SAMPLERS = 1000

    def __init__(self, model, system_shape, r, num_samples, num_flips, seed=0, chain_id0=0):
        self.args = dict(num_samples=num_samples, num_flips=num_flips, seed=seed, chain_id0=chain_id0)


def _vmc_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from qmcnn_b200 import vmc

        class M(object):
            r = 5
        a, b = vmc.make_samplers(M(), (6, 6), 64, 2500, 2, seed=9, sampler_cls=_RecordingSampler)
        out[rank] = (a.args, b.args)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_run_vmc_gives_every_rank_its_own_chains():
    """under torchrun every rank must own different global chains (chain_id0 = rank * num_samplers);
    identical ids would make the all-reduced batch world_size copies of the same samples."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_vmc_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][0]["chain_id0"] == 0 and out[1][0]["chain_id0"] == 64
    assert out[0][1]["chain_id0"] == 0 and out[1][1]["chain_id0"] == 1000     # capped at MAX_NUM_SAMPLERS chains
    assert out[0][0]["seed"] == out[1][0]["seed"] == 9 and out[0][1]["seed"] == 10
