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
