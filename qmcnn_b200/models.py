"""Wavefunction models with the reference's constructor and ``factors`` call
signatures (reference ``models.py``), backed by the CUDA forward kernel.

Parameters live in ONE flat fp32 CUDA tensor in the reference's variable
creation order and HWIO layout; ``model.params[name]`` are views into it
("filters", "bias_vis", "bias_hid" / "filters_%d", "bias_%d").
"""
import numpy as np
import torch

from . import _lib
from .helpers import scope_op


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class _Model(object):
    SCALE = 1E-2

    def _init_params(self, specs, device, seed):
        if not torch.cuda.is_available():
            raise _lib.QmcError("qmcnn_b200 needs a CUDA device (B200); there is no CPU fallback")
        _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        total = sum(int(np.prod(s)) for _, s in specs)
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(seed) if seed is not None else int(np.random.randint(2 ** 31)))
        self.flat = (torch.randn(total, generator=gen, dtype=torch.float32) * self.SCALE).to(self.device)
        self.names = [n for n, _ in specs]
        self.params, o = {}, 0
        for n, s in specs:
            size = int(np.prod(s))
            self.params[n] = self.flat[o:o + size].view(*s)
            o += size
        self._handles = {}
        # tuning / cross-check knobs for the handles this model creates (dict: flags = _lib.FLAG_* bits,
        # max_warps, ip_group, ip_chunks); defaults when empty.  Tests use it to run every decomposition.
        self.tuning = {}

    @property
    def num_params(self):
        return self.flat.numel()

    def set_flat_params(self, flat):
        self.flat.copy_(torch.as_tensor(flat, dtype=torch.float32).reshape(-1))

    def handle(self, system_shape):
        """The library handle for this model on an Ly x Lx lattice, parameters synced."""
        system_shape = tuple(int(s) for s in system_shape)
        if len(system_shape) != 2 or self.n_dims != 2:
            raise _lib.QmcError("this call is covered by the tuned 2-D kernels only (n_dims == 2); 1-D / 3-D "
                                "lattices have factors, Sampler and the energy estimators (qmc_nd_*)")
        key = (system_shape, tuple(sorted(self.tuning.items())))
        h = self._handles.get(key)
        if h is None:
            h = _lib.Handle(self._kind, self.k, self._channels, system_shape[0], system_shape[1],
                            self.device.index or 0, self.tuning)
            assert h.num_params == self.flat.numel()
            self._handles[key] = h
        _lib.check(h.ptr, _lib.load().qmc_set_params(h.ptr, self.flat.data_ptr(),
                                                     _stream_ptr(self.device)), "qmc_set_params")
        return h

    # ---- 1-D / 3-D lattices: the generic path (include/qmcnn_b200.h, qmc_nd_*) ------------
    def nd_desc(self, system_shape):
        system_shape = tuple(int(s) for s in system_shape)
        if len(system_shape) != self.n_dims:
            raise _lib.QmcError("system_shape %r does not have n_dims = %d axes" % (system_shape, self.n_dims))
        d = _lib.nd_desc(self._kind, self.k, self._channels, system_shape)
        if _lib.load().qmc_nd_num_params(d) != self.flat.numel():
            raise _lib.QmcError("nd: parameter count mismatch: %s" % _lib.load().qmc_nd_last_error().decode())
        return d

    def nd_scratch(self, desc, units):
        n = _lib.load().qmc_nd_scratch_floats(desc, self.device.index or 0, int(units))
        return torch.empty(max(n, 4), dtype=torch.float32, device=self.device)

    def nd_forward(self, spins, system_shape, want_factors=True, want_logpsi=False):
        """factors / log psi of UN-padded states (N, prod(shape)) on a 1-D / 3-D (or 2-D) lattice."""
        d = self.nd_desc(system_shape)
        n = int(np.prod(system_shape))
        spins = torch.as_tensor(spins, device=self.device).reshape(-1, n).to(torch.int8).contiguous()
        N = spins.shape[0]
        factors = torch.empty((N, n), dtype=torch.complex64, device=self.device) if want_factors else None
        logpsi = torch.empty(N, dtype=torch.complex64, device=self.device) if want_logpsi else None
        scratch = self.nd_scratch(d, N)
        _lib.check_nd(_lib.load().qmc_nd_forward(
            d, self.device.index or 0, self.flat.data_ptr(), spins.data_ptr(), N, scratch.data_ptr(),
            factors.data_ptr() if want_factors else None, logpsi.data_ptr() if want_logpsi else None,
            _stream_ptr(self.device)), "qmc_nd_forward")
        return factors, logpsi

    # ---- forward ------------------------------------------------------------
    def _halo(self):
        return (self.r - 1) // 2

    def forward_unpadded(self, spins, system_shape, want_factors=True, want_logpsi=False, cache=None):
        """K1 on UN-padded int8 spins (N, Ly*Lx). Returns (factors, logpsi, cache)."""
        h = self.handle(system_shape)
        spins = torch.as_tensor(spins, device=self.device).reshape(-1, h.n)
        if spins.dtype != torch.int8:
            spins = spins.to(torch.int8)
        spins = spins.contiguous()
        N = spins.shape[0]
        if cache is None:
            cache = torch.empty(N * h.cache_floats, dtype=torch.float32, device=self.device)
        factors = torch.empty((N, h.n), dtype=torch.complex64, device=self.device) if want_factors else None
        logpsi = torch.empty(N, dtype=torch.complex64, device=self.device) if want_logpsi else None
        _lib.check(h.ptr, _lib.load().qmc_logpsi_forward(
            h.ptr, spins.data_ptr(), N, cache.data_ptr(),
            factors.data_ptr() if want_factors else None,
            logpsi.data_ptr() if want_logpsi else None, _stream_ptr(self.device)), "qmc_logpsi_forward")
        return factors, logpsi, cache

    @scope_op("factors")
    def factors(self, x, periodic=False):
        """``models.py:31-67`` / ``95-131``: the reference's ``factors`` is a plain VALID
        cross-correlation stack, so it accepts ANY +-1 array at least ``r`` wide and returns
        complex64 ``(N,) + (shape - r + 1)`` - e.g. the ``(2K-1)^d`` windows of ``mcmc_tf.py:81-84``
        give ``K^d`` factors.

        The CUDA kernels evaluate periodic lattices.  A VALID output at centre c only reads the
        input inside c +- (r-1)/2, so the VALID result of ``x`` is the interior of the periodic
        factors of ``x`` itself taken as a lattice: that is what runs here (no host sync, no
        assumption about ``x``).  ``periodic=True``: the caller guarantees that ``x`` is a lattice
        wrap-padded by (r-1)//2 (what ``pad`` produces, ``sampler.py:85-88``); only the inner lattice
        is evaluated then - the same numbers for (1 + (r-1)/L)^d less work.  Nothing is checked.
        """
        x = torch.as_tensor(x, device=self.device)
        halo = self._halo()
        if x.dim() != self.n_dims + 1:
            raise _lib.QmcError("factors: expected (N,) + a %d-dimensional +-1 array" % self.n_dims)
        full = tuple(int(s) for s in x.shape[1:])
        inner = tuple(s - 2 * halo for s in full)
        if min(inner) < 1:
            raise _lib.QmcError("factors: the input must be at least r = %d wide along every axis" % self.r)
        core = (slice(None),) + tuple(slice(halo, halo + s) for s in inner)
        if periodic:
            lattice, shape, out_sl = x[core], inner, None
        else:
            lattice, shape, out_sl = x, full, core
        flat = lattice.reshape(x.shape[0], -1)
        if self.n_dims != 2:
            f, _ = self.nd_forward(flat, shape)
        else:
            f, _, _ = self.forward_unpadded(flat, shape)
        f = f.view((x.shape[0],) + shape)
        return f if out_sl is None else f[out_sl].contiguous()

    def log_psi(self, spins, system_shape):
        """log psi of UN-padded states (N, Ly*Lx) -> complex64 (N,)."""
        if self.n_dims != 2:
            return self.nd_forward(spins, system_shape, want_factors=False, want_logpsi=True)[1]
        return self.forward_unpadded(spins, system_shape, want_factors=False, want_logpsi=True)[1]


class CRBM(_Model):
    """Convolutional marginalised RBM, ``models.py:6-67``."""

    def __init__(self, k, pad_size, alpha, n_dims, device=None, seed=None):
        if n_dims not in (1, 2, 3):
            raise _lib.QmcError("n_dims must be 1, 2 or 3")
        self.k, self.pad_size, self.alpha, self.n_dims = k, pad_size, alpha, n_dims
        self.r = k
        self._kind, self._channels = _lib.MODEL_CRBM, [2 * alpha]
        self._init_params([("filters", (k,) * n_dims + (1, 2 * alpha)), ("bias_vis", (2,)),
                           ("bias_hid", (2 * alpha,))], device, seed)


class DCRBM(_Model):
    """Deep convolutional marginalised RBM, ``models.py:70-131``."""

    def __init__(self, k, layers, n_dims, device=None, seed=None):
        if n_dims not in (1, 2, 3):
            raise _lib.QmcError("n_dims must be 1, 2 or 3")
        self.k, self.layers, self.n_dims = k, list(layers), n_dims
        self.r = len(self.layers) * (k - 1) + 1
        self._kind, self._channels = _lib.MODEL_DCRBM, list(layers)
        chans = [1] + self.layers
        specs = []
        for l, (cin, cout) in enumerate(zip(chans, chans[1:])):
            specs += [("filters_%d" % l, (k,) * n_dims + (cin, cout)), ("bias_%d" % l, (cout,))]
        self._init_params(specs, device, seed)
