"""Metropolis sampler with the reference's ``Sampler`` interface (reference
``sampler.py``), backed by the persistent CUDA sweep kernel.

State is eager torch CUDA tensors owned by this object: int8 un-padded spins,
the opaque activation cache and the sample buffer.  The reference's variable
attributes (``current_samples_var`` etc.) are provided as properties in the
reference's shapes and dtypes.
"""
import numpy as np
import torch

from . import _lib
from .helpers import pad, scope_op
from .models import _stream_ptr


class Sampler(object):
    """``sampler.py:7-177``."""

    MAX_NUM_SAMPLERS = 1000
    SWEEPFACTOR = 10
    THERMFACTOR = 4

    def __init__(self, model, system_shape, r, num_samples, num_flips, seed=0, chain_id0=0):
        self.model = model
        self.system_shape = tuple(int(s) for s in system_shape)
        self.r = r
        self.num_samples = num_samples
        self.num_flips = num_flips
        self.n_dims = len(self.system_shape)
        self.num_spins = int(np.prod(self.system_shape))
        self.full_window_shape = (r * 2 - 1,) * self.n_dims
        self.full_window_size = int(np.prod(self.full_window_shape))
        self.half_window_shape = (r,) * self.n_dims
        self.half_window_size = int(np.prod(self.half_window_shape))
        # sampler.py:29-38
        self.num_samplers = min(num_samples, self.MAX_NUM_SAMPLERS)
        self.its_per_sample = self.num_spins * self.SWEEPFACTOR
        self.samples_per_sampler = num_samples // self.num_samplers
        self.therm_its = self.samples_per_sampler * self.its_per_sample * self.THERMFACTOR
        self.sample_its = (self.therm_its + (self.samples_per_sampler - 1) * self.its_per_sample + 1)
        self.padded_shape = tuple(s + r - 1 for s in self.system_shape)
        self.padded_size = int(np.prod(self.padded_shape))
        if r != model.r:
            raise _lib.QmcError("Sampler: r=%d does not match the model's receptive field %d" % (r, model.r))
        if self.num_samplers * self.samples_per_sampler != num_samples:
            raise _lib.QmcError("Sampler: num_samples=%d is not a multiple of num_samplers=%d "
                                "(the reference's final reshape fails too, sampler.py:176-177; "
                                "raise Sampler.MAX_NUM_SAMPLERS)" % (num_samples, self.num_samplers))
        self.new_samples = True                                    # sampler.py:40
        self.seed = int(seed)
        self.chain_id0 = int(chain_id0)      # global id of chain 0 (rank offset under data parallelism)
        self.device = model.device
        S, n = self.num_samplers, self.num_spins
        self._nd = getattr(model, "n_dims", 2) != 2       # 1-D / 3-D lattices: the generic path (qmc_nd_sweep)
        from .symmetry import SymmetrizedModel
        if not self._nd and not isinstance(model, SymmetrizedModel):
            # Outside the incremental kernels' coverage (flip bounding box + receptive field wider than the
            # lattice: two independent flip sites of a deep model, sampler.py:106-122) the chain runs the
            # reference's own algorithm - one full network evaluation per proposal - on the generic path.
            h = model.handle(self.system_shape)
            if _lib.load().qmc_sweep_workspace_floats(h.ptr, S, num_flips) == 0:
                self._nd = True
        if self._nd:
            self._sym = False
            self._desc = model.nd_desc(self.system_shape)
            self._spins = torch.zeros((S, n), dtype=torch.int8, device=self.device)
            self._factors = torch.zeros((S, n), dtype=torch.complex64, device=self.device)
            self._scratch = model.nd_scratch(self._desc, S)
            self._samples = torch.zeros((self.samples_per_sampler, S, n), dtype=torch.int8, device=self.device)
            self._n_accept = torch.zeros(1, dtype=torch.int64, device=self.device)
            self.flip_positions_var = self.accept_sample_var = self._initial_states = None
            self._step_base = 0
            self._resets = 0
            self.accept_trace = self.logratio_trace = None
            return
        self._h = model.handle(self.system_shape)
        lib = _lib.load()
        self._sym = isinstance(model, SymmetrizedModel)
        nimg = model.NSYM if self._sym else 1
        self._spins = torch.zeros((S, n), dtype=torch.int8, device=self.device)
        self._cache = torch.zeros(nimg * S * self._h.cache_floats, dtype=torch.float32, device=self.device)
        if self._sym:
            self._log_rel = torch.zeros((S, nimg, 2), dtype=torch.float64, device=self.device)
            ws = lib.qmc_sym_sweep_workspace_floats(self._h.ptr, S, num_flips, nimg)
        else:
            ws = lib.qmc_sweep_workspace_floats(self._h.ptr, S, num_flips)
        if ws == 0:
            raise _lib.QmcError("Sampler: this model / lattice / num_flips combination is outside the "
                                "incremental sweep's coverage (needs bounding box + r - 1 <= L)")
        self._workspace = torch.empty(ws, dtype=torch.float32, device=self.device)
        self._samples = torch.zeros((self.samples_per_sampler, S, n), dtype=torch.int8, device=self.device)
        self._n_accept = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.flip_positions_var = None      # fed-in proposals (parity mode) - sampler.py:60-64
        self.accept_sample_var = None       # fed-in uniforms                - sampler.py:65-69
        self._initial_states = None
        self._step_base = 0                 # Philox step offset: a fresh stream per mcmc_op call
        self._resets = 0                    # fresh-lattice draws so far (keys the initial-state stream)
        self.accept_trace = None
        self.logratio_trace = None

    # ---- reference-shaped views of the state ------------------------------------
    @property
    def current_samples_var(self):
        """int32 [S, padded_size], wrap-padded like sampler.py:44-48."""
        S = self.num_samplers
        g = self._spins.view((S,) + self.system_shape).to(torch.int32)
        return pad(g, self.system_shape, [(self.r - 1) // 2] * self.n_dims).reshape(S, -1)

    @property
    def current_factors_var(self):
        """complex64 [S, num_spins] (sampler.py:49-53), recomputed from the spins."""
        if self._nd:
            return self._factors
        m = self.model.base if self._sym else self.model
        return m.forward_unpadded(self._spins, self.system_shape)[0]

    @property
    def samples_var(self):
        return self._samples.to(torch.int32)

    @property
    def spins(self):
        """int8 [S, num_spins] current un-padded states."""
        return self._spins

    @property
    def acceptance_count(self):
        return int(self._n_accept.item())

    # ---- parity-mode inputs -----------------------------------------------------
    def feed(self, initial_states=None, flip_positions=None, accept_sample=None):
        """Feed the random draws of sampler.py:74-75 / 95-100 instead of generating them."""
        if initial_states is not None:
            self._initial_states = torch.as_tensor(initial_states, device=self.device) \
                .reshape(self.num_samplers, self.num_spins).to(torch.int8)
        if (flip_positions is None) != (accept_sample is None):
            raise _lib.QmcError("feed: give both flip_positions and accept_sample or neither")
        if flip_positions is not None:
            fp = torch.as_tensor(flip_positions, device=self.device) \
                .to(torch.int32).reshape(-1, self.num_samplers, self.num_flips).contiguous()
            ua = torch.as_tensor(accept_sample, device=self.device) \
                .to(torch.float32).reshape(-1, self.num_samplers).contiguous()
            # the kernels index shared and global memory with these sites: refuse anything off the lattice
            if fp.shape[0] != ua.shape[0]:
                raise _lib.QmcError("feed: flip_positions cover %d steps, accept_sample %d" % (fp.shape[0], ua.shape[0]))
            if fp.numel() and (int(fp.min()) < 0 or int(fp.max()) >= self.num_spins):
                raise _lib.QmcError("feed: flip_positions must lie in [0, %d)" % self.num_spins)
            self.flip_positions_var, self.accept_sample_var = fp, ua

    # ---- reference ops -------------------------------------------------------------
    @scope_op()
    def mcmc_reset(self):
        """``sampler.py:72-101``: fresh or persistent chains; factor cache refreshed
        under the CURRENT parameters; sample buffer zeroed."""
        if self.new_samples:
            if self._initial_states is not None:
                self._spins.copy_(self._initial_states)
            else:
                # iid +-1 (sampler.py:74-75), keyed by (seed, GLOBAL chain id, reset count): a chain's start does
                # not depend on how the chains are sharded over ranks
                rc = _lib.load().qmc_init_spins(self.device.index or 0, self._spins.data_ptr(), self.num_samplers,
                                                self.num_spins, self.seed, self.chain_id0, self._resets,
                                                _stream_ptr(self.device))
                if rc != 0:
                    raise _lib.QmcError("qmc_init_spins failed (%d): %s"
                                        % (rc, (_lib.load().qmc_last_error(None) or b"?").decode()))
                self._resets += 1
        if self._nd:
            self._factors = self.model.nd_forward(self._spins, self.system_shape)[0]      # sampler.py:85-88
            self._samples.zero_()
            return
        if self._sym:
            # caches of the 8 images + relative log amplitudes (per-site factor differences, double): 2 launches
            self.model.forward_images(self._spins, self.system_shape, caches=self._cache, log_rel=self._log_rel,
                                      want_logpsi=False)
        else:
            self.model.forward_unpadded(self._spins, self.system_shape, want_factors=False,
                                        cache=self._cache)
        self._samples.zero_()

    def _sweep(self, step0, n_steps, trace=False):
        if self._nd:
            return self._sweep_nd(step0, n_steps, trace)
        h = self.model.handle(self.system_shape)
        fed = self.flip_positions_var is not None
        fp = ua = None
        if fed:
            if step0 + n_steps > self.flip_positions_var.shape[0]:
                raise _lib.QmcError("fed-in proposals cover %d steps, need %d"
                                    % (self.flip_positions_var.shape[0], step0 + n_steps))
            fp = self.flip_positions_var[step0:step0 + n_steps]
            ua = self.accept_sample_var[step0:step0 + n_steps]
        S = self.num_samplers
        if trace:
            self.accept_trace = torch.zeros((n_steps, S), dtype=torch.uint8, device=self.device)
            self.logratio_trace = torch.zeros((n_steps, S), dtype=torch.float32, device=self.device)
        if self._sym:
            imgs = self.model.sync().contiguous()
            lib = _lib.load()
            _lib.check(h.ptr, lib.qmc_set_image_params(h.ptr, imgs.shape[0], imgs.data_ptr(),
                                                       _stream_ptr(self.device)), "qmc_set_image_params")
            _lib.check(h.ptr, lib.qmc_metropolis_sweep_sym(
                h.ptr, imgs.shape[0], self._spins.data_ptr(), self._cache.data_ptr(), self._log_rel.data_ptr(),
                self._workspace.data_ptr(), S, self.num_flips, step0 + (0 if fed else self._step_base), n_steps,
                fp.data_ptr() if fed else None, ua.data_ptr() if fed else None, self.seed, self.chain_id0,
                self.therm_its + (0 if fed else self._step_base), self.its_per_sample,
                self._samples.data_ptr(), self.samples_per_sampler,
                self.accept_trace.data_ptr() if trace else None,
                self.logratio_trace.data_ptr() if trace else None,
                self._n_accept.data_ptr(), _stream_ptr(self.device)), "qmc_metropolis_sweep_sym")
            return
        _lib.check(h.ptr, _lib.load().qmc_metropolis_sweep(
            h.ptr, self._spins.data_ptr(), self._cache.data_ptr(), self._workspace.data_ptr(),
            S, self.num_flips, step0 + (0 if fed else self._step_base), n_steps,
            fp.data_ptr() if fed else None, ua.data_ptr() if fed else None,
            self.seed, self.chain_id0,
            self.therm_its + (0 if fed else self._step_base), self.its_per_sample,
            self._samples.data_ptr(), self.samples_per_sampler,
            self.accept_trace.data_ptr() if trace else None,
            self.logratio_trace.data_ptr() if trace else None,
            self._n_accept.data_ptr(), _stream_ptr(self.device)), "qmc_metropolis_sweep")

    def _sweep_nd(self, step0, n_steps, trace):
        fed = self.flip_positions_var is not None
        fp = ua = None
        if fed:
            if step0 + n_steps > self.flip_positions_var.shape[0]:
                raise _lib.QmcError("fed-in proposals cover %d steps, need %d"
                                    % (self.flip_positions_var.shape[0], step0 + n_steps))
            fp = self.flip_positions_var[step0:step0 + n_steps]
            ua = self.accept_sample_var[step0:step0 + n_steps]
        S = self.num_samplers
        if trace:
            self.accept_trace = torch.zeros((n_steps, S), dtype=torch.uint8, device=self.device)
            self.logratio_trace = torch.zeros((n_steps, S), dtype=torch.float32, device=self.device)
        _lib.check_nd(_lib.load().qmc_nd_sweep(
            self._desc, self.device.index or 0, self.model.flat.data_ptr(), self._spins.data_ptr(),
            self._factors.data_ptr(), self._scratch.data_ptr(), S, self.num_flips,
            step0 + (0 if fed else self._step_base), n_steps,
            fp.data_ptr() if fed else None, ua.data_ptr() if fed else None, self.seed, self.chain_id0,
            self.therm_its + (0 if fed else self._step_base), self.its_per_sample,
            self._samples.data_ptr(), self.samples_per_sampler,
            self.accept_trace.data_ptr() if trace else None, self.logratio_trace.data_ptr() if trace else None,
            self._n_accept.data_ptr(), _stream_ptr(self.device)), "qmc_nd_sweep")

    @scope_op()
    def mcmc_step(self, i, trace=False):
        """``sampler.py:104-155``: one Metropolis step of every chain; returns i + 1."""
        self._sweep(int(i), 1, trace)
        return i + 1

    @scope_op()
    def mcmc_op(self, n_its=None, trace=False):
        """``sampler.py:158-177``: reset, run the whole loop in ONE kernel launch, return
        int32 (num_samples, num_spins), row = j * num_samplers + chain."""
        self.mcmc_reset()
        n_its = self.sample_its if n_its is None else int(n_its)
        self._sweep(0, n_its, trace)
        if self.flip_positions_var is None:
            self._step_base += max(n_its, self.sample_its)      # never reuse a part of the Philox stream
        return self._samples.view(self.num_samples, self.num_spins).to(torch.int32)

    def samples_int8(self):
        """The sample buffer without the int32 widening (what the energy kernel consumes)."""
        return self._samples.view(self.num_samples, self.num_spins)
