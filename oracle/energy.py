"""Oracle restatement of the local-energy estimators, VMC loss, gradient and
the TF-1 Adam update (reference ``mcmc_tf.py:35-179``).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

The reference reads ``K, H, SYSTEM_SHAPE, NUM_SPINS`` from module globals
(``mcmc_tf.py:15-25``); here they are explicit arguments.  ``K`` is the
receptive field ``r`` of the model (``k`` for CRBM, ``D(k-1)+1`` for DCRBM).
"""
import numpy as np

from .helpers import pad, all_windows, interactions


def _cdtype(model):
    return np.complex64 if model.dtype == np.float32 else np.complex128


def ising_energy(model, states, system_shape, K, H=1.0):
    """TFIM local energy per spin by the window trick, ``mcmc_tf.py:59-90``."""
    states = np.asarray(states, np.int32)
    n = states.shape[0]
    n_dims = len(system_shape)
    num_spins = int(np.prod(system_shape))
    full, half = (2 * K - 1,) * n_dims, (K,) * n_dims
    padded = pad(states.reshape((n,) + tuple(system_shape)), system_shape,
                 [(K - 1) // 2] * n_dims)                               # :72-73
    factors = model.factors(padded).reshape(n, -1)                      # :74
    factor_windows = all_windows(factors, system_shape, half)           # :75
    spin_windows = all_windows(states, system_shape, full)              # :76
    flipper = np.ones(int(np.prod(full)), np.int32)
    flipper[(flipper.size - 1) // 2] = -1                               # :77-78
    flipped = spin_windows * flipper                                    # :79
    factors_flipped = model.factors(
        flipped.reshape((n * num_spins,) + full)).reshape(n, num_spins, -1)  # :80-83
    log_pop = (factors_flipped - factor_windows).sum(2)                 # :85
    aligned = interactions(states, system_shape).sum((1, 2))            # :86
    energy = (-model.dtype(H) * np.exp(log_pop).sum(1)
              - aligned.astype(_cdtype(model)))                         # :87-88
    return (energy / num_spins).astype(_cdtype(model))                  # :89


def heisenberg_energy(model, states, system_shape, K):
    """Marshall-signed AFM Heisenberg local energy per spin,
    ``mcmc_tf.py:93-141``."""
    states = np.asarray(states, np.int32)
    n = states.shape[0]
    n_dims = len(system_shape)
    num_spins = int(np.prod(system_shape))
    full, half = (2 * K + 1,) * n_dims, (K + 2,) * n_dims               # :105-108
    padded = pad(states.reshape((n,) + tuple(system_shape)), system_shape,
                 [(K - 1) // 2] * n_dims)
    factors = model.factors(padded).reshape(n, -1)
    factor_windows = all_windows(factors, system_shape, half)           # :114
    spin_windows = all_windows(states, system_shape, full)              # :115
    flippers = np.ones((n_dims,) + full, np.int32)
    centre = tuple((s - 1) // 2 for s in full)                          # :120
    for d in range(n_dims):                                             # :118-125
        nb = tuple(c + 1 if i == d else c for i, c in enumerate(centre))
        flippers[(d,) + centre] = -1
        flippers[(d,) + nb] = -1
    flippers = flippers.reshape(n_dims, -1)
    flipped = spin_windows[:, None, :, :] * flippers[None, :, None, :]  # :128
    factors_flipped = model.factors(
        flipped.reshape((n * n_dims * num_spins,) + full)).reshape(
            n, n_dims, num_spins, -1)                                   # :129-132
    log_pop = (factors_flipped - factor_windows[:, None, :, :]).sum(3)  # :134
    ints = interactions(states, system_shape).astype(_cdtype(model))    # :135
    terms = -(1 - ints) * np.exp(log_pop) + ints                        # :137
    return (terms.sum((1, 2)) / num_spins).astype(_cdtype(model))       # :138-140


def batched_op(fn, states, batch_size):
    """``mcmc_tf.py:144-153``: map ``fn`` over chunks of ``batch_size`` rows."""
    states = np.asarray(states)
    assert states.shape[0] % batch_size == 0
    return np.concatenate([fn(states[i:i + batch_size])
                           for i in range(0, states.shape[0], batch_size)])


def loss_op(factors, energies):
    """Covariance loss Re[<E conj(log psi)> - <E><conj(log psi)>],
    ``mcmc_tf.py:35-56``.  factors: (N, ...) complex; energies: (N,)."""
    factors = np.asarray(factors)
    n = factors.shape[0]
    energies = np.asarray(energies).astype(factors.dtype)
    log_psi_conj = np.conj(factors.reshape(n, -1).sum(1))               # :51-52
    e_avg = energies.sum() / n                                          # :53
    loss = (energies * log_psi_conj).sum() / n - e_avg * log_psi_conj.sum() / n
    return np.real(loss)                                                # :54-56


def vmc_gradient(model, samples_padded, energies, dtype=np.float64):
    """d loss_op / d params by torch autograd on CPU - the stand-in for TF
    autodiff (``mcmc_tf.py:172-177``).  Independent of the hand-written CUDA
    backward.  Returns the flat gradient in ``model.flat_params()`` order.
    """
    import torch
    import torch.nn.functional as F
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    names = model.names
    p = {n: torch.tensor(np.asarray(model.params[n], dtype=dtype), dtype=tdt,
                         requires_grad=True) for n in names}
    x = torch.tensor(np.asarray(samples_padded), dtype=tdt)[:, None]   # N, C, *spatial
    nd = model.n_dims
    convnd = (F.conv1d, F.conv2d, F.conv3d)[nd - 1]
    bshape = (1, -1) + (1,) * nd

    def conv(h, w):            # [*k, I, O] -> [O, I, *k]; F.convNd is a cross-correlation
        return convnd(h, w.permute([nd + 1, nd] + list(range(nd))))

    if hasattr(model, "alpha"):
        a = model.alpha
        theta = conv(x, p["filters"]) + p["bias_hid"].reshape(bshape)
        theta = torch.complex(theta[:, :a], theta[:, a:])
        act = torch.log(torch.exp(theta) + torch.exp(-theta)).sum(1)
        ps = model.pad_size
        xu = x[(slice(None), 0) + tuple(slice(ps, x.shape[2 + d] - ps) for d in range(nd))]
        factors = act + torch.complex(p["bias_vis"][0] * xu, p["bias_vis"][1] * xu)
    else:
        h = x
        D = len(model.layers)
        for l in range(D):
            h = conv(h, p["filters_%d" % l]) + p["bias_%d" % l].reshape(bshape)
            if l != D - 1:
                h = torch.tanh(h)
        sep = model.layers[-1] // 2
        theta = torch.complex(h[:, :sep], h[:, sep:])
        factors = torch.log(torch.exp(theta) + torch.exp(-theta)).sum(1)
    n = factors.shape[0]
    e = torch.tensor(np.asarray(energies), dtype=factors.dtype)
    lpc = torch.conj(factors.reshape(n, -1).sum(1))
    loss = ((e * lpc).sum() / n - (e.sum() / n) * lpc.sum() / n).real
    loss.backward()
    return np.concatenate([p[n].grad.numpy().ravel() for n in names]), float(loss.detach())


def adam_tf1_step(p, g, m, v, t, lr=3e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """One ``tf.train.AdamOptimizer`` update (``mcmc_tf.py:176``), TF-1 form:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps), t = 1, 2, ...
    Returns (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v
