"""Oracle restatement of the lattice symmetry group (``symmetry.ipynb`` cell 0).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

G = D4 x| T on an M x M torus as integer affine 3x3 matrices acting on
(i, j, 1); the translation column is reduced mod M.  ``np.int`` and
``stack(<generator>)`` in the notebook are the only things changed (they no
longer exist in numpy 2).
"""
from itertools import product as _iproduct

import numpy as np
from numpy.linalg import matrix_power


def mod(g, M):
    g = np.array(g, dtype=np.int64, copy=True)
    g[..., :, 2] = g[..., :, 2] % M
    # the homogeneous row stays (0, 0, 1): 1 % M == 1 for M > 1
    return g


def product(A, B, M):
    return mod(np.stack([a @ b for a, b in _iproduct(A, B)], 0), M)


def generate(a, n, M):
    return mod(np.stack([matrix_power(a, i) for i in range(n)], 0), M)


def inv(g):
    return np.rint(np.linalg.inv(g)).astype(np.int64)


R0 = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]])     # 90 degree rotation
M0 = np.array([[1, 0, 0], [0, -1, 0], [0, 0, 1]])     # mirror j -> -j
I0 = np.array([[1, 0, 1], [0, 1, 0], [0, 0, 1]])      # unit translation in i
J0 = np.array([[1, 0, 0], [0, 1, 1], [0, 0, 1]])      # unit translation in j


def d4(M):
    """8 point-group elements, index = 2*rot + mirror."""
    return product(generate(R0, 4, M), generate(M0, 2, M), M)


def translations(M):
    """M*M translations, index = M*a + b."""
    return product(generate(I0, M, M), generate(J0, M, M), M)


def group(M):
    """All 8*M*M elements, element = D4[p] . T[t], index = M*M*p + t."""
    return product(d4(M), translations(M), M)


def plot(g, M):
    """The notebook's ``plot``: grid with ``grid[g.x mod M] = label(x)``."""
    n = M * M
    coords = np.stack(np.unravel_index(np.arange(n), (M, M))
                      + (np.ones(n, dtype=np.int64),), 1)
    t = np.einsum("ab,nb->na", g, coords)
    res = np.full((M, M), -1, dtype=np.int64)
    res[t[:, 0] % M, t[:, 1] % M] = np.arange(n)
    return res


def site_permutation(g, M):
    """perm with (g.s)[y] = s[perm[y]] (flat indices), i.e. (g.s)(y) = s(g^-1 y)."""
    return plot(g, M).ravel()


def neighbours(grid, M):
    flat = grid.ravel()
    n = M * M
    coords = np.stack(np.unravel_index(np.arange(n), (M, M)), 1)
    offs = np.array([[0, 1], [0, -1], [1, 0], [-1, 0]])
    nc = (coords[:, None, :] + offs) % M
    nb = flat[nc[..., 0] * M + nc[..., 1]]
    return {int(flat[i]): sorted(int(v) for v in nb[i]) for i in range(n)}


def d4_filter_images(filters):
    """The 8 images W o g of an odd, centred HWIO filter such that
    psi(g.s; W) = psi(s; W o g).  Index order matches ``d4``."""
    out = []
    k = filters.shape[0]
    c = (k - 1) // 2
    for g in d4(3 * k):                      # any M > k: only the linear part matters
        lin = g[:2, :2]
        img = np.empty_like(filters)
        for u in range(k):
            for v in range(k):
                su, sv = lin @ np.array([u - c, v - c])
                img[u, v] = filters[su + c, sv + c]
        out.append(img)
    return out


def symmetrised_log_psi(model, states, system_shape, full_group=False):
    """log[(1/|G|) sum_g psi(g.s)] for states (N, L*L) on an L x L torus.
    ``full_group`` averages all 8 L^2 elements (slow; the check), otherwise the
    8 point-group images (translations are a no-op for these models)."""
    from .helpers import pad
    L = system_shape[0]
    elems = group(L) if full_group else d4(L)
    states = np.asarray(states)
    halo = (model.r - 1) // 2
    logs = []
    for g in elems:
        perm = site_permutation(g, L)
        gs = states[:, perm].reshape((-1,) + tuple(system_shape))
        logs.append(model.log_psi(pad(gs, system_shape, [halo, halo])))
    logs = np.stack(logs, 0)
    m = logs.real.max(0)
    return np.log(np.exp(logs - m).mean(0)) + m


# ---------------------------------------------------------------------------------------
# Symmetry-averaged amplitude (BASELINE config 4).  The reference stops at the group; the
# amplitude is DEFINED in SURVEY.md section 8: psi_sym(s) = (1/|G|) sum_g psi(g.s).  Because
# log psi is translation invariant this equals the mean over the 8 point-group images, and
# psi(g.s; W) = psi(s; W o g) with every layer's filter transformed by g.
# ---------------------------------------------------------------------------------------
def image_models(model):
    """The 8 models psi(.; W o g), g in D4, sharing the base parameters (biases unchanged)."""
    out = []
    for p in range(8):
        m = model.astype(model.dtype)
        m.params = dict(model.params)
        for name, arr in model.params.items():
            if name.startswith("filters"):
                m.params[name] = d4_filter_images(arr)[p]
        out.append(m)
    return out


def log_psi_images(model, states, system_shape):
    """(8, N) complex log psi of the 8 images for un-padded states (N, L*L)."""
    from .helpers import pad
    halo = (model.r - 1) // 2
    x = pad(np.asarray(states).reshape((-1,) + tuple(system_shape)), system_shape, [halo, halo])
    return np.stack([m.log_psi(x) for m in image_models(model)], 0)


def log_mean_exp(logs):
    m = logs.real.max(0)
    return np.log(np.exp(logs - m).mean(0)) + m


def sym_local_energy(model, states, system_shape, hamiltonian, H=1.0):
    """E_loc of psi_sym per spin by brute force over connected configurations (float64 truth):
    TFIM: [-H sum_i psi_sym(s^i)/psi_sym(s) - sum_<ij> s_i s_j] / n
    Heisenberg (Marshall): sum_<ij> [s_i s_j == 1 ? 1 : -1 - 2 psi_sym(s^ij)/psi_sym(s)] / n"""
    states = np.asarray(states)
    L0, L1 = system_shape
    n = L0 * L1
    base = log_mean_exp(log_psi_images(model, states, system_shape))
    idx = np.arange(n).reshape(system_shape)
    e = np.zeros(states.shape[0], dtype=np.complex128)
    for d in range(2):
        nb = np.roll(idx, -1, d).ravel()
        for i in range(n):
            sisj = states[:, i] * states[:, nb[i]]
            if hamiltonian == "tfim":
                e -= sisj
            else:
                fl = states.copy(); fl[:, i] *= -1; fl[:, nb[i]] *= -1
                ratio = np.exp(log_mean_exp(log_psi_images(model, fl, system_shape)) - base)
                e += np.where(sisj == 1, 1.0, -1.0 - 2.0 * ratio)
    if hamiltonian == "tfim":
        for i in range(n):
            fl = states.copy(); fl[:, i] *= -1
            e -= H * np.exp(log_mean_exp(log_psi_images(model, fl, system_shape)) - base)
    return e / n


class SymSampler(object):
    """Metropolis sampling of |psi_sym|^2: sampler.py:104-155 with psi replaced by psi_sym,
    full recompute of all 8 images every step (the slow reference-style algorithm)."""

    def __init__(self, model, system_shape, num_flips):
        self.model, self.system_shape, self.num_flips = model, tuple(system_shape), num_flips
        self.n = int(np.prod(system_shape))

    def reset(self, initial_states):
        self.states = np.asarray(initial_states, np.int32).reshape(-1, self.n).copy()
        self.logs = log_psi_images(self.model, self.states, self.system_shape)

    def step(self, centers, u, force_mask=None):
        S = self.states.shape[0]
        prop = self.states.copy()
        for f in range(self.num_flips):
            prop[np.arange(S), centers[:, f]] *= -1
        new_logs = log_psi_images(self.model, prop, self.system_shape)
        log_ratio = log_mean_exp(new_logs) - log_mean_exp(self.logs)
        self.last_log_ratio = log_ratio
        mask = np.abs(np.exp(log_ratio)) ** 2 > u
        self.last_own_mask = mask
        if force_mask is not None:
            mask = np.asarray(force_mask, bool)
        self.states[mask] = prop[mask]
        self.logs[:, mask] = new_logs[:, mask]
        return mask
