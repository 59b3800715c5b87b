"""Oracle restatement of the lattice/window helpers (reference ``helpers.py``).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  numpy in, numpy out;
every function works for any number of lattice dimensions like the reference.
"""
import numpy as np


def create_index_matrix(data_shape, window_shape):
    """Wrapped flat indices of the window centred on every lattice site.

    Follows ``helpers.py:8-34``: row ``i`` lists, in row-major window order, the
    flat index (periodic wrap) of ``site_i - (w-1)//2 + box`` (``:28-32``).
    Returns int32 ``(n_sites, n_window)``.
    """
    data_shape = tuple(int(s) for s in data_shape)
    window_shape = tuple(int(w) for w in window_shape)
    n_dims = len(data_shape)
    sites = np.stack(np.unravel_index(np.arange(int(np.prod(data_shape))),
                                      data_shape), 1)            # (n, d)
    box = np.stack(np.unravel_index(np.arange(int(np.prod(window_shape))),
                                    window_shape), 1)             # (w, d)
    offset = (np.array(window_shape) - 1) // 2
    coords = sites[:, None, :] - offset[None, None, :] + box[None, :, :]
    coords %= np.array(data_shape)[None, None, :]
    flat = np.zeros(coords.shape[:2], dtype=np.int64)
    for d in range(n_dims):
        flat = flat * data_shape[d] + coords[..., d]
    return flat.astype(np.int32)


def unpad(x, pad_size):
    """Strip ``pad_size[d]`` entries from both ends of every lattice axis.

    ``helpers.py:52-70``; axis 0 is the batch axis.
    """
    x = np.asarray(x)
    sl = (slice(None),) + tuple(
        slice(p, x.shape[d + 1] - p) for d, p in enumerate(pad_size))
    return x[sl]


def pad(x, system_shape, pad_size):
    """Periodic halo: tile x3 along every lattice axis, then slice
    ``s - p`` off each side (``helpers.py:73-91``)."""
    x = np.asarray(x)
    tiled = np.tile(x, (1,) + (3,) * len(pad_size))
    return unpad(tiled, tuple(s - p for s, p in zip(system_shape, pad_size)))


def all_windows(x, system_shape, window_shape):
    """Every window of every row: ``(N, n_sites) -> (N, n_sites, n_window)``
    (``helpers.py:149-168``)."""
    x = np.asarray(x)
    return x[:, create_index_matrix(system_shape, window_shape)]


def gather_windows(x, centers, system_shape, window_shape):
    """Per-row window around a per-row flat centre (``helpers.py:94-118``;
    dead code in the reference, kept for API parity). ``x`` is ``(N, n_sites)``,
    ``centers`` ``(N,)`` flat indices -> ``(N, n_window)``."""
    x = np.asarray(x)
    idx = create_index_matrix(system_shape, window_shape)[np.asarray(centers)]
    return np.take_along_axis(x, idx.astype(np.int64), axis=1)


def update_windows(x, centers, updates, mask, system_shape, window_shape):
    """Masked scatter of ``updates`` into the windows around ``centers``
    (``helpers.py:121-146``). Returns the updated copy of ``x``."""
    x = np.array(x, copy=True)
    idx = create_index_matrix(system_shape, window_shape)[np.asarray(centers)]
    upd = np.asarray(updates).reshape(x.shape[0], -1)
    for n in np.nonzero(np.asarray(mask))[0]:
        x[n, idx[n]] = upd[n]
    return x


def interactions(states, system_shape):
    """s_i * s_{i+e_d} for every axis d, periodic (``helpers.py:171-195``:
    neighbour = ``np.roll(indices, -1, d)``). ``(N, n_sites) -> (N, d, n_sites)``."""
    states = np.asarray(states)
    n = int(np.prod(system_shape))
    indices = np.arange(n).reshape(system_shape)
    out = [states * states[:, np.roll(indices, -1, d).ravel()]
           for d in range(len(system_shape))]
    return np.stack(out, 1)
