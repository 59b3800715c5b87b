"""Lattice / window helpers with the reference's names and conventions
(reference ``helpers.py``), as eager torch functions on any device.

Inside the CUDA kernels the periodic halo and the windows are index
arithmetic; these functions exist so user scripts that import them keep
working.  They are plumbing, not the hot path.
"""
import functools

import numpy as np
import torch


def create_index_matrix(data_shape, window_shape):
    """``helpers.py:8-34``: wrapped flat indices of the window centred (offset
    ``(w-1)//2``) on every site, row-major window order. int32 (n_sites, n_window)."""
    data_shape = tuple(int(s) for s in data_shape)
    window_shape = tuple(int(w) for w in window_shape)
    grids = np.meshgrid(*[np.arange(s) for s in data_shape], indexing="ij")
    sites = np.stack([g.ravel() for g in grids], 1)
    wg = np.meshgrid(*[np.arange(w) - (w - 1) // 2 for w in window_shape], indexing="ij")
    box = np.stack([g.ravel() for g in wg], 1)
    coords = (sites[:, None, :] + box[None, :, :]) % np.array(data_shape)
    strides = np.cumprod((1,) + data_shape[:0:-1])[::-1]
    return (coords * strides).sum(-1).astype(np.int32)


def scope_op(name=None):
    """``helpers.py:37-49`` wrapped graph builders in a TF name scope; here it
    opens an NVTX range when CUDA is available."""
    def decorator(function):
        @functools.wraps(function)
        def wrapper(*args, **kwargs):
            if torch.cuda.is_available():
                torch.cuda.nvtx.range_push(name or function.__name__)
                try:
                    return function(*args, **kwargs)
                finally:
                    torch.cuda.nvtx.range_pop()
            return function(*args, **kwargs)
        return wrapper
    return decorator


def unpad(x, pad_size):
    """``helpers.py:52-70``."""
    sl = (slice(None),) + tuple(slice(p, x.shape[d + 1] - p) for d, p in enumerate(pad_size))
    return x[sl]


def pad(x, system_shape, pad_size):
    """``helpers.py:73-91``: periodic halo (== ``np.pad(mode='wrap')``)."""
    for d, p in enumerate(pad_size):
        if p:
            idx = (torch.arange(-p, system_shape[d] + p, device=x.device) % system_shape[d])
            x = x.index_select(d + 1, idx)
    return x


@functools.lru_cache(maxsize=64)
def _index_matrix_cached(data_shape, window_shape):
    return create_index_matrix(data_shape, window_shape)


def _im(x, system_shape, window_shape):
    im = _index_matrix_cached(tuple(system_shape), tuple(window_shape))
    return torch.as_tensor(im, device=x.device, dtype=torch.long)


def all_windows(x, system_shape, window_shape):
    """``helpers.py:149-168``: (N, n_sites) -> (N, n_sites, n_window)."""
    return x[:, _im(x, system_shape, window_shape)]


def gather_windows(x, centers, system_shape, window_shape):
    """``helpers.py:94-118``: per-row window around a per-row flat centre."""
    idx = _im(x, system_shape, window_shape)[centers.long()]
    return torch.gather(x, 1, idx)


def update_windows(x, centers, updates, mask, system_shape, window_shape):
    """``helpers.py:121-146``: masked scatter; returns the updated tensor (in place)."""
    idx = _im(x, system_shape, window_shape)[centers.long()]
    rows = torch.nonzero(mask).ravel()
    upd = updates.reshape(x.shape[0], -1)
    x[rows[:, None], idx[rows]] = upd[rows]
    return x


def interactions(states, system_shape):
    """``helpers.py:171-195``: s_i * s_{i+e_d} per axis, (N, n_dims, n_sites)."""
    n = int(np.prod(system_shape))
    indices = np.arange(n).reshape(system_shape)
    out = []
    for d in range(len(system_shape)):
        nb = torch.as_tensor(np.roll(indices, -1, d).ravel(), device=states.device, dtype=torch.long)
        out.append(states * states[:, nb])
    return torch.stack(out, 1)
