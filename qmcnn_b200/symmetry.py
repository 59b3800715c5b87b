"""Lattice symmetry group and the symmetry-averaged amplitude (reference ``symmetry.ipynb``).

The notebook builds G = D4 x| T on an M x M torus as integer affine 3x3 matrices and checks
the group axioms; it defines no amplitude.  Following SURVEY.md section 8, the amplitude is

    psi_sym(s) = (1/|G|) sum_g psi(g.s)

Circular-conv models are translation invariant, so the average collapses to the 8 point-group
images, and psi(g.s; W) = psi(s; W o g) with every layer's filter rotated / mirrored.  The
CUDA path therefore evaluates 8 *parameter images* of the same model on the same lattice:

    log psi_sym = logsumexp_g log psi_g - log 8
    E_loc[psi_sym](s) = sum_g p_g(s) E_loc[psi_g](s),   p_g = psi_g(s) / sum_g' psi_g'(s)
    d log psi_sym / dW = sum_g p_g  d log psi_g / dW   (pulled back through the tap permutation)

and the Metropolis sweep of |psi_sym|^2 runs in ``qmc_metropolis_sweep_sym``.
"""
from itertools import product as _iproduct

import numpy as np
import torch

from . import _lib
from .models import _stream_ptr

# ---- the notebook's group (numpy, host side) ----------------------------------------------
R0 = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]])     # 90 degree rotation
M0 = np.array([[1, 0, 0], [0, -1, 0], [0, 0, 1]])     # mirror j -> -j
I0 = np.array([[1, 0, 1], [0, 1, 0], [0, 0, 1]])      # unit translations
J0 = np.array([[1, 0, 0], [0, 1, 1], [0, 0, 1]])


def mod(g, M):
    g = np.array(g, dtype=np.int64, copy=True)
    g[..., :, 2] = g[..., :, 2] % M
    return g


def product(A, B, M):
    return mod(np.stack([a @ b for a, b in _iproduct(A, B)], 0), M)


def generate(a, n, M):
    return mod(np.stack([np.linalg.matrix_power(a, i) for i in range(n)], 0), M)


def d4(M):
    """8 point-group elements, index = 2*rot + mirror."""
    return product(generate(R0, 4, M), generate(M0, 2, M), M)


def translations(M):
    return product(generate(I0, M, M), generate(J0, M, M), M)


def group(M):
    return product(d4(M), translations(M), M)


def plot(g, M):
    """The notebook's ``plot``: grid[g.x mod M] = label(x)."""
    n = M * M
    coords = np.stack(np.unravel_index(np.arange(n), (M, M)) + (np.ones(n, dtype=np.int64),), 1)
    t = np.einsum("ab,nb->na", g, coords)
    res = np.full((M, M), -1, dtype=np.int64)
    res[t[:, 0] % M, t[:, 1] % M] = np.arange(n)
    return res


def neighbours(grid, M):
    flat = grid.ravel()
    coords = np.stack(np.unravel_index(np.arange(M * M), (M, M)), 1)
    offs = np.array([[0, 1], [0, -1], [1, 0], [-1, 0]])
    nc = (coords[:, None, :] + offs) % M
    nb = flat[nc[..., 0] * M + nc[..., 1]]
    return {int(flat[i]): sorted(int(v) for v in nb[i]) for i in range(M * M)}


def d4_tap_permutations(k):
    """perm[p][u*k+v] = flat tap index of W that image p reads at tap (u, v):
    (W o g_p)[u, v] = W[g_p.(u-c, v-c) + c]."""
    c = (k - 1) // 2
    out = np.zeros((8, k * k), dtype=np.int64)
    for p, g in enumerate(d4(3 * k)):
        lin = g[:2, :2]
        for u in range(k):
            for v in range(k):
                su, sv = lin @ np.array([u - c, v - c])
                out[p, u * k + v] = (su + c) * k + (sv + c)
    return out


# ---- the symmetrised model on the GPU -------------------------------------------------------
class SymmetrizedModel(object):
    """psi_sym built from ``base`` (a CRBM or DCRBM).  ``base.flat`` stays the one trainable
    parameter vector; the 8 image vectors are gathers of it."""

    NSYM = 8

    def __init__(self, base):
        self.base = base
        self.device = base.device
        self.r = base.r
        self.k = base.k
        self.names, self.params, self.flat = base.names, base.params, base.flat
        taps = torch.as_tensor(d4_tap_permutations(base.k), device=self.device)
        # gather index [8, P]: image_flat[g] = base.flat[index[g]]
        idx = torch.arange(base.flat.numel(), device=self.device).repeat(self.NSYM, 1)
        o = 0
        for n in base.names:
            shape = base.params[n].shape
            size = base.params[n].numel()
            if n.startswith("filters"):
                inner = size // (base.k * base.k)             # C_in * C_out per tap
                blk = torch.arange(size, device=self.device).view(base.k * base.k, inner)
                for g in range(self.NSYM):
                    idx[g, o:o + size] = o + blk[taps[g]].reshape(-1)
            o += size
        self.index = idx
        self._images = [type(base).__new__(type(base)) for _ in range(self.NSYM)]
        for g, im in enumerate(self._images):
            im.__dict__.update({k: v for k, v in base.__dict__.items() if k not in ("flat", "params", "_handles")})
            im.flat = torch.empty_like(base.flat)
            im.params = {}
            im._handles = {}
        self.sync()

    @property
    def num_params(self):
        return self.base.flat.numel()

    def sync(self):
        """Refresh the 8 image parameter vectors from ``base.flat``."""
        imgs = self.base.flat[self.index]                      # [8, P]
        for g, im in enumerate(self._images):
            im.flat.copy_(imgs[g])
        return imgs

    def images(self):
        self.sync()
        return self._images

    def handle(self, system_shape):
        return self.base.handle(system_shape)

    # ---- the images as an extra grid dimension of the same launches (qmc_*_sym) -------------
    def _bind(self, system_shape):
        """Handle with the current image parameter blocks uploaded (qmc_set_image_params)."""
        h = self.base.handle(system_shape)
        imgs = self.sync().contiguous()
        _lib.check(h.ptr, _lib.load().qmc_set_image_params(h.ptr, self.NSYM, imgs.data_ptr(), _stream_ptr(self.device)),
                   "qmc_set_image_params")
        return h

    def _states(self, states, h):
        st = torch.as_tensor(states, device=self.device).reshape(-1, h.n)
        if st.dtype != torch.int8:
            st = st.to(torch.int8)
        return st.contiguous()

    def forward_images(self, spins, system_shape, caches=None, log_rel=None, want_logpsi=True):
        """One launch for the 8 forwards + one for the combination: fills ``caches`` [8, N, cache_floats]
        and ``log_rel`` [N, 8, 2] (float64, log psi_g - log psi_0) when given; returns log psi_sym (N,)."""
        h = self._bind(system_shape)
        st = self._states(spins, h)
        N = st.shape[0]
        if caches is None:
            caches = torch.empty(self.NSYM * N * h.cache_floats, dtype=torch.float32, device=self.device)
        out = torch.empty(N, dtype=torch.complex64, device=self.device) if want_logpsi else None
        _lib.check(h.ptr, _lib.load().qmc_logpsi_forward_sym(
            h.ptr, self.NSYM, st.data_ptr(), N, caches.data_ptr(),
            log_rel.data_ptr() if log_rel is not None else None,
            out.data_ptr() if want_logpsi else None, _stream_ptr(self.device)), "qmc_logpsi_forward_sym")
        return out

    def log_psi(self, spins, system_shape):
        """log psi_sym = log((1/8) sum_g psi_g), complex64 (N,)."""
        return self.forward_images(spins, system_shape)

    def local_energy(self, hamiltonian, h_field, states, system_shape, moments=None):
        """E_loc[psi_sym] = sum_g p_g E_loc[psi_g] in three launches (forward of all images, connected
        configurations of all images, combination) - ``qmc_local_energy_sym``."""
        h = self._bind(system_shape)
        st = self._states(states, h)
        N = st.shape[0]
        out = torch.empty(N, dtype=torch.complex64, device=self.device)
        if N == 0:
            return out
        lib = _lib.load()
        ws = torch.empty(lib.qmc_sym_energy_workspace_floats(h.ptr, self.NSYM, N), dtype=torch.float32, device=self.device)
        _lib.check(h.ptr, lib.qmc_local_energy_sym(
            h.ptr, self.NSYM, hamiltonian, float(h_field), st.data_ptr(), N, ws.data_ptr(), out.data_ptr(),
            moments.data_ptr() if moments is not None else None, _stream_ptr(self.device)), "qmc_local_energy_sym")
        return out

    def gradient(self, states, weights, system_shape):
        """sum_n Re[w_n conj(d log psi_sym,n / dp)] in ``base.flat`` order: the 8 image gradients come out of
        one launch sequence (``qmc_logpsi_backward_sym``) and are scattered back through the tap permutation."""
        h = self._bind(system_shape)
        st = self._states(states, h)
        N = st.shape[0]
        out = torch.zeros_like(self.base.flat)
        if N == 0:
            return out
        lib = _lib.load()
        w = weights.to(torch.complex64).contiguous()
        gi = torch.zeros((self.NSYM, self.num_params), dtype=torch.float32, device=self.device)
        ws = torch.empty(lib.qmc_sym_backward_workspace_floats(h.ptr, self.NSYM, N), dtype=torch.float32, device=self.device)
        _lib.check(h.ptr, lib.qmc_logpsi_backward_sym(h.ptr, self.NSYM, st.data_ptr(), w.data_ptr(), N, ws.data_ptr(),
                                                      gi.data_ptr(), _stream_ptr(self.device)), "qmc_logpsi_backward_sym")
        out.index_add_(0, self.index.reshape(-1), gi.reshape(-1))
        return out

    # ---- image-by-image composition (8 x the plain entry points): the cross-check of the launches above ----
    def log_psi_images(self, spins, system_shape):
        """(8, N) complex64."""
        return torch.stack([im.log_psi(spins, system_shape) for im in self.images()], 0)

    def image_weights(self, logs):
        """p_g = psi_g / sum psi_g from (8, N) log amplitudes (complex128 for the softmax)."""
        z = logs.to(torch.complex128)
        z = z - z.real.max(0, keepdim=True).values
        w = torch.exp(z)
        return w / w.sum(0, keepdim=True)

    def log_psi_composed(self, spins, system_shape):
        logs = self.log_psi_images(spins, system_shape).to(torch.complex128)
        m = logs.real.max(0).values
        return (torch.log(torch.exp(logs - m).mean(0)) + m).to(torch.complex64)

    def local_energy_composed(self, energy_fn, states, system_shape, **kw):
        """sum_g p_g E_loc[psi_g]; ``energy_fn`` is ising_energy / heisenberg_energy."""
        logs = self.log_psi_images(states, system_shape)
        p = self.image_weights(logs)
        e = torch.stack([energy_fn(im, states, system_shape=system_shape, **kw) for im in self._images], 0)
        return (p * e.to(torch.complex128)).sum(0).to(torch.complex64)

    def gradient_composed(self, grad_fn, states, weights, system_shape):
        logs = self.log_psi_images(states, system_shape)
        p = self.image_weights(logs)
        out = torch.zeros_like(self.base.flat)
        for g, im in enumerate(self._images):
            gi = grad_fn(im, states, (weights.to(torch.complex128) * torch.conj(p[g])).to(torch.complex64),
                         system_shape)
            out.index_add_(0, self.index[g], gi)
        return out
