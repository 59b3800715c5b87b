"""Shared helpers for the GPU parity tests: build the same model in the CUDA
package and in the oracle, with identical parameters."""
import numpy as np
import torch

import oracle
import qmcnn_b200 as q


def make_pair(kind, L, scale, seed, dtype=np.float32, **kw):
    """Returns (cuda_model, oracle_model) with identical parameters ~N(0, scale)."""
    rng = np.random.default_rng(seed)
    if kind == "crbm":
        k, alpha = kw.get("k", 5), kw.get("alpha", 4)
        om = oracle.CRBM(k, (k - 1) // 2, alpha, 2, rng=rng, scale=scale, dtype=dtype)
        gm = q.CRBM(k, (k - 1) // 2, alpha, 2, seed=seed)
    else:
        k, layers = kw.get("k", 3), kw["layers"]
        om = oracle.DCRBM(k, layers, 2, rng=rng, scale=scale, dtype=dtype)
        gm = q.DCRBM(k, layers, 2, seed=seed)
    gm.set_flat_params(om.flat_params().astype(np.float32))
    for n in om.names:   # the views must agree with the oracle's named arrays
        assert np.array_equal(gm.params[n].cpu().numpy(), om.params[n].astype(np.float32)), n
    return gm, om


def rand_states(rng, n, shape):
    return (rng.integers(0, 2, (n, int(np.prod(shape)))) * 2 - 1).astype(np.int32)


def padded(om, states, shape):
    halo = (om.r - 1) // 2
    return oracle.pad(states.reshape((-1,) + tuple(shape)), shape, [halo, halo])


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
