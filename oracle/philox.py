"""Philox-4x32-10 counter RNG in numpy (Salmon et al., SC'11; Random123).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

The reference draws its proposals with ``tf.random_uniform``
(``sampler.py:95-100``), which cannot be reproduced; the CUDA sweep instead
defines its own stream and this file is its CPU twin:

    counter = (step_lo, step_hi, chain_lo, chain_hi), key = (seed_lo, seed_hi)
    words r0..r3 = philox4x32_10(counter, key)
    flip position f (f < num_flips <= 3) = mulhi32(r_f, num_spins)
    uniform u in [0,1)                   = (r3 >> 8) * 2**-24

Fresh initial lattices (``sampler.py:74-79`` draws them with ``tf.random_uniform``) come from the
same generator on counters the proposal stream never uses (bit 31 of the second word set):

    counter = (block, 0x80000000 | reset_index, chain_lo, chain_hi);  spin 128*block + 32*w + b
    = +1 if bit b of word r_w is set else -1
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter)[..., i].astype(np.uint64) for i in range(4)]
    k0 = np.asarray(key)[..., 0].astype(np.uint64)
    k1 = np.asarray(key)[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & _MASK,
             (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & _MASK]
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack(c, -1).astype(np.uint32)


def sweep_randoms(seed, chain_ids, step0, n_steps, num_flips, num_spins):
    """The (flip_positions, accept_sample) arrays the CUDA sweep generates
    in-kernel for steps ``step0 .. step0+n_steps-1`` of the given global chains.

    Returns int32 (n_steps, S, num_flips) and float32 (n_steps, S), i.e. the
    shapes of ``flip_positions_var`` / ``accept_sample_var`` (``sampler.py:60-69``).
    """
    assert 1 <= num_flips <= 3
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    steps = (np.uint64(step0) + np.arange(n_steps, dtype=np.uint64))
    ctr = np.zeros((n_steps, chain_ids.size, 4), dtype=np.uint32)
    ctr[..., 0] = (steps & _MASK)[:, None]
    ctr[..., 1] = (steps >> np.uint64(32))[:, None]
    ctr[..., 2] = (chain_ids & _MASK)[None, :]
    ctr[..., 3] = (chain_ids >> np.uint64(32))[None, :]
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    pos = ((r[..., :num_flips].astype(np.uint64) * np.uint64(num_spins))
           >> np.uint64(32)).astype(np.int32)
    u = ((r[..., 3] >> np.uint32(8)).astype(np.float32)
         * np.float32(2.0 ** -24))
    return pos, u


def initial_spins(seed, chain_ids, num_spins, reset_index=0):
    """The +-1 lattices ``qmc_init_spins`` writes for the given global chains (int32 (S, num_spins)):
    a function of (seed, global chain id, reset_index) only, so independent of how chains are sharded."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    nblk = (num_spins + 127) // 128
    ctr = np.zeros((chain_ids.size, nblk, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.uint32(0x80000000 | (int(reset_index) & 0x7FFFFFFF))
    ctr[..., 2] = (chain_ids & _MASK)[:, None]
    ctr[..., 3] = (chain_ids >> np.uint64(32))[:, None]
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)                                       # (S, nblk, 4)
    bits = (r[..., None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)   # (S, nblk, 4, 32)
    return (bits.reshape(chain_ids.size, -1)[:, :num_spins].astype(np.int32) * 2 - 1)
