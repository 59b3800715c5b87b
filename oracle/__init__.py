"""CPU oracle for the qmcnn variational-Monte-Carlo hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``qmcnn_b200`` (the product) may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker or as the timed CPU arm.

PARITY UNPINNED: the reference's arithmetic lives entirely in TensorFlow-1.x
(un-vendored, un-pinned; ``models.py:57-67``, ``sampler.py:74-155``,
``mcmc_tf.py:49-141``) which is not installed here and cannot be installed
(no wheel, no network); the reference ships no tests or golden vectors.  This
oracle is therefore a numpy restatement of the cited reference lines.  It is
pinned by (see ``tests/test_oracle_*.py``): pad == np.pad('wrap'), window trick
== brute-force flip + full forward, translation invariance of log psi, the D4
filter-fold identity, the notebook's group-axiom assertions, the sampler
bookkeeping table, dense-Hamiltonian local energies on a 3x3 lattice and
exact-diagonalisation ground states, Philox-4x32-10 known-answer vectors, and
torch-autograd / finite differences of ``loss_op`` for the gradient.
"""
from .helpers import (create_index_matrix, pad, unpad, all_windows,
                      gather_windows, update_windows, interactions)
from .models import CRBM, DCRBM
from .sampler import Sampler
from .energy import (ising_energy, heisenberg_energy, loss_op, batched_op,
                     vmc_gradient, adam_tf1_step)
from . import symmetry, philox

__all__ = [
    "create_index_matrix", "pad", "unpad", "all_windows", "gather_windows",
    "update_windows", "interactions", "CRBM", "DCRBM", "Sampler",
    "ising_energy", "heisenberg_energy", "loss_op", "batched_op",
    "vmc_gradient", "adam_tf1_step", "symmetry", "philox",
]
