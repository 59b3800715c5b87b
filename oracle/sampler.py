"""Oracle restatement of the Metropolis sampler (reference ``sampler.py``).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

All randomness (initial lattices, proposals, acceptance uniforms) is an INPUT:
the reference draws it with ``tf.random_uniform`` (``sampler.py:74-75,95-100``),
which is not reproducible, so parity is defined on identical fed-in arrays.
Every step re-evaluates the full network on the flipped state exactly like
``sampler.py:117-122`` - this is deliberately the slow reference algorithm.
"""
import numpy as np

from .helpers import pad, unpad


class Sampler(object):
    MAX_NUM_SAMPLERS = 1000        # sampler.py:10
    SWEEPFACTOR = 10               # sampler.py:11
    THERMFACTOR = 4                # sampler.py:12

    def __init__(self, model, system_shape, r, num_samples, num_flips):
        self.model = model
        self.system_shape = tuple(system_shape)
        self.r = r
        self.num_samples = num_samples
        self.num_flips = num_flips
        self.n_dims = len(self.system_shape)
        self.num_spins = int(np.prod(self.system_shape))
        # bookkeeping, sampler.py:29-38
        self.num_samplers = min(num_samples, self.MAX_NUM_SAMPLERS)
        self.its_per_sample = self.num_spins * self.SWEEPFACTOR
        self.samples_per_sampler = num_samples // self.num_samplers
        self.therm_its = (self.samples_per_sampler * self.its_per_sample
                          * self.THERMFACTOR)
        self.sample_its = (self.therm_its + (self.samples_per_sampler - 1)
                           * self.its_per_sample + 1)
        self.padded_shape = tuple(s + r - 1 for s in self.system_shape)
        self.padded_size = int(np.prod(self.padded_shape))
        self.halo = (r - 1) // 2

        self.new_samples = True                                     # :40
        S = self.num_samplers
        cdt = np.complex64 if model.dtype == np.float32 else np.complex128
        self.current_samples = np.zeros((S, self.padded_size), np.int32)   # :44-48
        self.current_factors = np.zeros((S, self.num_spins), cdt)          # :49-53
        self.samples = np.zeros((self.samples_per_sampler, S, self.num_spins),
                                np.int32)                                   # :54-59
        self.flip_positions = None                                          # :60-64
        self.accept_sample = None                                           # :65-69
        # flipper table, sampler.py:108-113: -1 at every wrap image of site c
        flipper = np.ones((self.num_spins, self.num_spins), np.int32)
        flipper[np.arange(self.num_spins), np.arange(self.num_spins)] = -1
        flipper = flipper.reshape((self.num_spins,) + self.system_shape)
        width = [[0, 0]] + [[self.halo, self.halo]] * self.n_dims
        self._flipper_padded = np.pad(flipper, width, "wrap").reshape(
            self.num_spins, self.padded_size)
        # diagnostics for the parity harness (not in the reference)
        self.last_log_ratio = None
        self.last_mask = None

    # ------------------------------------------------------------------
    def mcmc_reset(self, initial_states, flip_positions, accept_sample):
        """``sampler.py:72-101`` with the three random draws fed in.

        initial_states: (S,)+system_shape +-1 ints, used only when
        ``self.new_samples`` (``:81-83``); flip_positions int32
        (n_its, S, num_flips); accept_sample float32 (n_its, S).
        """
        S = self.num_samplers
        if self.new_samples:
            fresh = pad(np.asarray(initial_states, np.int32), self.system_shape,
                        [self.halo] * self.n_dims)                  # :76-77
            states = fresh.reshape(S, -1)                            # :78-79
        else:
            states = self.current_samples
        factors = self.model.factors(
            states.reshape((S,) + self.padded_shape)).reshape(S, -1)   # :85-88
        self.current_samples = states.astype(np.int32).copy()
        self.current_factors = factors
        self.samples[...] = 0                                        # :93-94
        self.flip_positions = np.asarray(flip_positions, np.int32)
        self.accept_sample = np.asarray(accept_sample, np.float32)

    def mcmc_step(self, i, force_mask=None):
        """``sampler.py:104-155``.  ``force_mask`` (parity-harness hook, not in the
        reference) overrides the accept decisions so that two implementations can
        be compared in lock-step after a tie."""
        S = self.num_samplers
        centers = self.flip_positions[i]                             # :106
        combined = np.prod(self._flipper_padded[centers], 1)         # :114-115
        flipped = self.current_samples * combined                    # :117
        flipped_factors = self.model.factors(
            flipped.reshape((S,) + self.padded_shape)).reshape(S, self.num_spins)
        log_ratio = (flipped_factors - self.current_factors).sum(1)  # :124 inner
        accept_prob = np.abs(np.exp(log_ratio)) ** 2                 # :123-124
        mask = accept_prob > self.accept_sample[i]                   # :125 strict
        self.last_own_mask = mask
        if force_mask is not None:
            mask = np.asarray(force_mask, bool)
        self.current_samples[mask] = flipped[mask]                   # :128-130
        self.current_factors[mask] = flipped_factors[mask]           # :131-133
        self.last_log_ratio, self.last_mask = log_ratio, mask
        k = i - self.therm_its
        if k >= 0 and k % self.its_per_sample == 0:                  # :148-152
            j = k // self.its_per_sample                             # :137
            self.samples[j] = self.unpadded_current()                # :138-145
        return i + 1

    def unpadded_current(self):
        S = self.num_samplers
        return unpad(self.current_samples.reshape((S,) + self.padded_shape),
                     (self.halo,) * self.n_dims).reshape(S, self.num_spins)

    def mcmc_op(self, initial_states, flip_positions, accept_sample, n_its=None):
        """``sampler.py:158-177``; ``n_its`` truncates the loop for tests."""
        self.mcmc_reset(initial_states, flip_positions, accept_sample)
        n_its = self.sample_its if n_its is None else n_its
        for i in range(n_its):
            self.mcmc_step(i)
        return self.samples.reshape(self.num_samples, self.num_spins)   # :176-177
