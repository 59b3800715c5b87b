"""Oracle restatement of the wavefunction models (reference ``models.py``).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

``dtype=np.float32`` mimics TensorFlow's float32/complex64 arithmetic (the
literal ``log(exp(t) + exp(-t))`` of ``models.py:65,130``); ``np.float64`` is the
ground-truth mode.  Parameters are plain numpy arrays in the reference's
variable names and HWIO layout; ``flat_params`` concatenates them in the
reference's creation order, which is the order the CUDA library takes them in.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

from .helpers import unpad


def _valid_xcorr(x, filters):
    """'VALID' cross-correlation, channels-last, no kernel flip (what
    ``tf.nn.conv{1,2,3}d`` computes, ``models.py:57-61``).

    x: (N, *spatial, C_in); filters: (*k, C_in, C_out) -> (N, *spatial-k+1, C_out)
    """
    n_dims = filters.ndim - 2
    kshape = filters.shape[:n_dims]
    win = sliding_window_view(x, kshape, axis=tuple(range(1, n_dims + 1)))
    # win: (N, *out, C_in, *k) -> contract (C_in, *k) with filters (*k, C_in, C_out)
    win = np.moveaxis(win, n_dims + 1, -1)            # (N, *out, *k, C_in)
    out_shape = win.shape[:n_dims + 1]
    a = np.ascontiguousarray(win).reshape(int(np.prod(out_shape)), -1)
    b = filters.reshape(-1, filters.shape[-1])
    return (a @ b).reshape(out_shape + (filters.shape[-1],))


def _log2cosh(theta):
    """Literal ``tf.log(tf.exp(theta) + tf.exp(-theta))``, principal branch
    per element (``models.py:65,130``)."""
    return np.log(np.exp(theta) + np.exp(-theta))


class _Base(object):
    SCALE = 1e-2

    def _cdtype(self):
        return np.complex64 if self.dtype == np.float32 else np.complex128

    @property
    def names(self):
        return list(self._names)

    def flat_params(self):
        """All parameters, reference creation order, C-order flattened."""
        return np.concatenate([self.params[n].ravel() for n in self._names])

    def set_flat_params(self, flat):
        flat = np.asarray(flat, dtype=self.dtype)
        o = 0
        for n in self._names:
            size = self.params[n].size
            self.params[n] = flat[o:o + size].reshape(self.params[n].shape).copy()
            o += size
        assert o == flat.size

    def astype(self, dtype):
        """Same parameters, different working precision."""
        other = object.__new__(type(self))
        other.__dict__.update(self.__dict__)
        other.dtype = np.dtype(dtype).type
        other.params = {k: v.astype(dtype) for k, v in self.params.items()}
        return other

    def log_psi(self, x):
        f = self.factors(x)
        return f.reshape(f.shape[0], -1).sum(1)


class CRBM(_Base):
    """``models.py:6-67``: one VALID conv (k^d x 1 x 2*alpha), hidden bias,
    complex theta = first alpha channels + i * last alpha, sum_ch log 2cosh,
    plus the complex visible bias times the centre spin."""

    def __init__(self, k, pad_size, alpha, n_dims, rng=None, scale=None,
                 dtype=np.float32):
        self.k, self.pad_size, self.alpha, self.n_dims = k, pad_size, alpha, n_dims
        self.dtype = np.dtype(dtype).type
        self.r = k
        rng = np.random.default_rng(0) if rng is None else rng
        scale = self.SCALE if scale is None else scale
        self._names = ["filters", "bias_vis", "bias_hid"]      # models.py:19-28
        shapes = {"filters": (k,) * n_dims + (1, 2 * alpha),
                  "bias_vis": (2,), "bias_hid": (2 * alpha,)}
        self.params = {n: (scale * rng.standard_normal(shapes[n])).astype(dtype)
                       for n in self._names}

    def factors(self, x):
        """x: (N,) + padded lattice, +-1 ints -> complex (N,) + (padded - k + 1)."""
        p = self.params
        xf = np.asarray(x).astype(self.dtype)                          # :51
        x_unpad = unpad(xf, (self.pad_size,) * self.n_dims)            # :52
        theta = _valid_xcorr(xf[..., None], p["filters"]) + p["bias_hid"]   # :59,62
        theta = (theta[..., :self.alpha]
                 + 1j * theta[..., self.alpha:]).astype(self._cdtype())     # :64
        act = _log2cosh(theta)                                         # :65
        bias = (p["bias_vis"][0] * x_unpad
                + 1j * (p["bias_vis"][1] * x_unpad)).astype(self._cdtype())  # :66
        return act.sum(-1) + bias                                      # :67


class DCRBM(_Base):
    """``models.py:70-131``: D VALID conv layers, tanh between, last layer split
    at ``layers[-1]//2`` into Re/Im, sum_ch log 2cosh.  No visible bias."""

    def __init__(self, k, layers, n_dims, rng=None, scale=None, dtype=np.float32):
        self.k, self.layers, self.n_dims = k, list(layers), n_dims
        self.dtype = np.dtype(dtype).type
        self.r = len(self.layers) * (k - 1) + 1
        rng = np.random.default_rng(0) if rng is None else rng
        scale = self.SCALE if scale is None else scale
        chans = [1] + self.layers                                      # :78
        self._names, self.params = [], {}
        for l, (cin, cout) in enumerate(zip(chans, chans[1:])):        # :85-92
            for name, shape in (("filters_%d" % l, (k,) * n_dims + (cin, cout)),
                                ("bias_%d" % l, (cout,))):
                self._names.append(name)
                self.params[name] = (scale * rng.standard_normal(shape)).astype(dtype)

    def activations(self, x):
        """All layer outputs (post-tanh for l < D, raw for the last)."""
        h = np.asarray(x).astype(self.dtype)[..., None]                # :110
        outs = []
        for l in range(len(self.layers)):                              # :113-126
            h = _valid_xcorr(h, self.params["filters_%d" % l]) \
                + self.params["bias_%d" % l]
            if l != len(self.layers) - 1:
                h = np.tanh(h)
            outs.append(h)
        return outs

    def factors(self, x):
        h = self.activations(x)[-1]
        sep = self.layers[-1] // 2                                     # :128
        theta = (h[..., :sep] + 1j * h[..., sep:]).astype(self._cdtype())   # :129
        return _log2cosh(theta).sum(-1)                                # :130-131
