"""CPU oracle for the qmcnn variational-Monte-Carlo hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``qmcnn_b200`` (the product) may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker or as the timed CPU arm.

PARITY PINNED TO THE REFERENCE'S OWN PYTHON (not to TensorFlow's kernels): the
reference's arithmetic lives in TensorFlow-1.x (un-vendored, un-pinned;
``models.py:57-67``, ``sampler.py:74-155``, ``mcmc_tf.py:49-141``), which cannot be
installed here, and the reference ships no tests or golden vectors.  So
``oracle/tf1_shim`` provides the 55 ``tf.*`` symbols the reference uses as eager
torch-CPU ops, ``tests/golden/make_golden.py`` imports /root/reference's
helpers.py / models.py / sampler.py / mcmc_tf.py UNMODIFIED over that shim and
records what they compute (Sampler.mcmc_op chains with every accept decision,
model.factors in 1/2/3-D, both energy estimators, loss_op gradients, two Adam
iterations, every helper), and ``tests/test_golden_oracle.py`` holds this numpy
restatement to those vectors: integers bit-exact, float64 1e-10, float32 1e-5.
What stays unpinned is TensorFlow's own kernel rounding and RNG streams (each
shim op restates the documented TF-1 semantics).  Independent pins on top
(``tests/test_oracle_pins.py``): pad == np.pad('wrap'), window trick == brute-force
flip + full forward, translation invariance, the D4 filter-fold identity, the
notebook's group axioms, dense-Hamiltonian local energies on 3x3 and exact
diagonalisation, Philox-4x32-10 known answers, finite differences of ``loss_op``.
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
