"""The torch-CPU timing restatement must agree with the numpy oracle."""
import numpy as np
import torch

import oracle
from oracle import torch_ref


def test_torch_ref_matches_numpy_oracle():
    rng = np.random.default_rng(3)
    L = 8
    for om in (oracle.CRBM(5, 2, 3, 2, rng=rng, scale=0.2), oracle.DCRBM(3, [4, 6, 4], 2, rng=rng, scale=0.3)):
        tm = torch_ref.TorchModel(om)
        s = (rng.integers(0, 2, (5, L, L)) * 2 - 1).astype(np.int32)
        halo = (om.r - 1) // 2
        want = om.factors(oracle.pad(s, (L, L), [halo, halo]))
        got = tm.factors(torch_ref._pad(torch.as_tensor(s), halo)).numpy()
        assert np.abs(got - want).max() < 1e-5
        e = torch_ref.ising_energy(tm, torch.as_tensor(s), H=0.7).numpy()
        ew = oracle.ising_energy(om, s.reshape(5, -1), (L, L), om.r, H=0.7)
        assert np.abs(e - ew).max() < 1e-4 * np.abs(ew).max()
        n_its = 30
        pos = rng.integers(0, L * L, (n_its, 5, 1)).astype(np.int32)
        u = rng.random((n_its, 5)).astype(np.float32)
        cur, _ = torch_ref.metropolis_steps(tm, torch.as_tensor(s), torch.as_tensor(pos), torch.as_tensor(u))
        smp = oracle.Sampler(om, (L, L), om.r, 5, 1)
        smp.mcmc_reset(s, pos, u)
        for i in range(n_its):
            smp.mcmc_step(i)
        assert np.array_equal(cur.numpy().reshape(5, -1).astype(np.int32), smp.unpadded_current())
