"""qmcnn_b200 - B200-native (sm_100a) drop-in for the variational-Monte-Carlo
hot path of dmaloneynygc/qmcnn: CRBM/DCRBM log-psi forward, batched Metropolis
sweep, TFIM/Heisenberg local energies and the log-psi gradient.

Same call surface as the reference's ``models.py`` / ``sampler.py`` /
``helpers.py`` / ``mcmc_tf.py``; eager torch CUDA tensors instead of TF graph
nodes; all arithmetic in hand-written CUDA kernels behind the C ABI of
``include/qmcnn_b200.h``.  No CPU fallback.
"""
from ._lib import QmcError, load as load_library, LIB_PATH
from ._lib import (FLAG_GENERIC_CONV, FLAG_SWEEP_CLASSIC, FLAG_SWEEP_INPLACE, FLAG_IP_FREE_RUNNING,
                   FLAG_ENERGY_CLASSIC, FLAG_ENERGY_INPLACE, FLAG_BACKWARD_GENERIC, FLAG_IP_ROWMAJOR_SITES,
                   FLAG_FORWARD_BLOCKED, FLAG_BACKWARD_SMEM)
from .helpers import (create_index_matrix, scope_op, pad, unpad, all_windows, gather_windows,
                      update_windows, interactions)
from .models import CRBM, DCRBM
from .sampler import Sampler
from . import distributed, symmetry
from .symmetry import SymmetrizedModel
from .vmc import run_vmc
from .mcmc import (ising_energy, heisenberg_energy, batched_op, loss_op, optimize_op, eval_op,
                   logpsi_gradient, AdamTF1, OptimizeStep)

__all__ = ["QmcError", "load_library", "LIB_PATH", "create_index_matrix", "scope_op", "pad", "unpad",
           "all_windows", "gather_windows", "update_windows", "interactions", "CRBM", "DCRBM",
           "Sampler", "SymmetrizedModel", "symmetry", "distributed", "ising_energy", "heisenberg_energy", "batched_op", "loss_op", "optimize_op",
           "eval_op", "logpsi_gradient", "AdamTF1", "OptimizeStep", "run_vmc"]
