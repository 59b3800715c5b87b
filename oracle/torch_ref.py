"""The reference ALGORITHM restated with torch CPU ops, for timing only.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Used by ``bench.py``'s
``cpu_baseline`` leg and ``--impl reference`` arm (TensorFlow itself cannot be
installed here): per Metropolis step the FULL network is re-evaluated on every
chain (``sampler.py:117-133``), the energy runs the network on all N*L^2
windows (``mcmc_tf.py:72-90``) - no incremental trick.  ``F.conv2d`` is the
multi-threaded oneDNN convolution, the closest stand-in for TF's CPU conv.
Checked against the numpy oracle in ``tests/test_oracle_torch_ref.py``.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .helpers import create_index_matrix


class TorchModel(object):
    def __init__(self, oracle_model):
        m = oracle_model
        self.is_crbm = hasattr(m, "alpha")
        self.r = m.r
        self.k = m.k
        t = lambda a: torch.tensor(np.asarray(a, np.float32))
        if self.is_crbm:
            self.alpha, self.pad_size = m.alpha, m.pad_size
            self.w = [t(m.params["filters"]).permute(3, 2, 0, 1).contiguous()]
            self.b = [t(m.params["bias_hid"])]
            self.bias_vis = t(m.params["bias_vis"])
        else:
            self.layers = m.layers
            self.w = [t(m.params["filters_%d" % l]).permute(3, 2, 0, 1).contiguous()
                      for l in range(len(m.layers))]
            self.b = [t(m.params["bias_%d" % l]) for l in range(len(m.layers))]

    def factors(self, x):
        """x: (N, H, W) padded +-1 -> complex64 (N, H-r+1, W-r+1); models.py:51-67 / 110-131."""
        h = x.to(torch.float32)[:, None]
        D = len(self.w)
        for l in range(D):
            h = F.conv2d(h, self.w[l], self.b[l])
            if l != D - 1:
                h = torch.tanh(h)
        half = h.shape[1] // 2
        theta = torch.complex(h[:, :half], h[:, half:])
        f = torch.log(torch.exp(theta) + torch.exp(-theta)).sum(1)
        if self.is_crbm:
            p = self.pad_size
            xu = x[:, p:x.shape[1] - p, p:x.shape[2] - p].to(torch.float32)
            f = f + torch.complex(self.bias_vis[0] * xu, self.bias_vis[1] * xu)
        return f


def _pad(x, p):
    return F.pad(x[:, None].to(torch.float32), (p, p, p, p), mode="circular")[:, 0]


def metropolis_steps(tm, states, flip_positions, accept_sample):
    """sampler.py:104-133 for n_its = flip_positions.shape[0] steps (num_flips = 1 or 2).
    states: (S, Ly, Lx) float/int tensor, modified copy returned with the accept count."""
    S, Ly, Lx = states.shape
    p = (tm.r - 1) // 2
    cur = states.clone().to(torch.float32)
    cur_f = tm.factors(_pad(cur, p)).reshape(S, -1)
    rows = torch.arange(S)
    n_acc = 0
    for i in range(flip_positions.shape[0]):
        prop = cur.reshape(S, -1).clone()
        for f in range(flip_positions.shape[2]):
            prop[rows, flip_positions[i, :, f].long()] *= -1
        prop = prop.reshape(S, Ly, Lx)
        new_f = tm.factors(_pad(prop, p)).reshape(S, -1)
        prob = torch.abs(torch.exp((new_f - cur_f).sum(1))) ** 2
        mask = prob > accept_sample[i]
        cur[mask] = prop[mask]
        cur_f[mask] = new_f[mask]
        n_acc += int(mask.sum())
    return cur, n_acc


def ising_energy(tm, states, H=1.0):
    """mcmc_tf.py:59-90 (window trick through the full network). states (N, Ly, Lx)."""
    N, Ly, Lx = states.shape
    n, K = Ly * Lx, tm.r
    flat = states.reshape(N, n).to(torch.float32)
    factors = tm.factors(_pad(states, (K - 1) // 2)).reshape(N, n)
    fw = factors[:, torch.as_tensor(create_index_matrix((Ly, Lx), (K, K)), dtype=torch.long)]
    sw = flat[:, torch.as_tensor(create_index_matrix((Ly, Lx), (2 * K - 1,) * 2), dtype=torch.long)].clone()
    sw[:, :, ((2 * K - 1) ** 2 - 1) // 2] *= -1
    ff = tm.factors(sw.reshape(N * n, 2 * K - 1, 2 * K - 1)).reshape(N, n, K * K)
    log_pop = (ff - fw).sum(2)
    g = states.to(torch.float32)
    aligned = (g * torch.roll(g, -1, 1)).sum((1, 2)) + (g * torch.roll(g, -1, 2)).sum((1, 2))
    return (-H * torch.exp(log_pop).sum(1) - aligned) / n
