"""Local-energy estimators, VMC loss and optimisation step with the reference's
names (reference ``mcmc_tf.py:35-194``), backed by the CUDA energy and backward
kernels.

The reference reads ``K, H, SYSTEM_SHAPE`` from module globals
(``mcmc_tf.py:15-25``); the same names exist here as module attributes with the
same defaults and can be overridden per call with keyword arguments.
"""
import numpy as np
import torch

from . import _lib
from .helpers import scope_op
from .models import _stream_ptr

LEARNING_RATE = 3E-3
K = 5
SYSTEM_SHAPE = (10, 10)
H = 1.0
ENERGY_BATCH_SIZE = 1000


def _prep(model, states, system_shape):
    system_shape = tuple(SYSTEM_SHAPE if system_shape is None else system_shape)
    states = torch.as_tensor(states, device=model.device)
    n = int(np.prod(system_shape))
    states = states.reshape(-1, n)
    if states.dtype != torch.int8:
        states = states.to(torch.int8)
    return states.contiguous(), system_shape, model.handle(system_shape)


def _is_sym(model):
    from .symmetry import SymmetrizedModel
    return isinstance(model, SymmetrizedModel)


def _local_energy(model, states, system_shape, hamiltonian, h_field, moments=None):
    if _is_sym(model):     # E_loc[psi_sym] = sum_g p_g E_loc[psi_g]
        shape = tuple(SYSTEM_SHAPE if system_shape is None else system_shape)
        return model.local_energy(hamiltonian, h_field, states, shape, moments)
    if getattr(model, "n_dims", 2) != 2:
        # 1-D / 3-D lattices: the generic path (one full network evaluation per connected configuration)
        shape = tuple(SYSTEM_SHAPE if system_shape is None else system_shape)
        d = model.nd_desc(shape)
        st = torch.as_tensor(states, device=model.device).reshape(-1, int(np.prod(shape))).to(torch.int8).contiguous()
        out = torch.empty(st.shape[0], dtype=torch.complex64, device=model.device)
        if st.shape[0]:
            scratch = model.nd_scratch(d, st.shape[0])
            _lib.check_nd(_lib.load().qmc_nd_local_energy(
                d, model.device.index or 0, hamiltonian, float(h_field), model.flat.data_ptr(), st.data_ptr(),
                st.shape[0], scratch.data_ptr(), out.data_ptr(), _stream_ptr(model.device)), "qmc_nd_local_energy")
        if moments is not None:
            moments += torch.stack([torch.tensor(float(out.numel()), device=out.device, dtype=torch.float64),
                                    out.real.double().sum(), out.imag.double().sum(),
                                    (out.real.double() ** 2 + out.imag.double() ** 2).sum()])
        return out
    states, system_shape, h = _prep(model, states, system_shape)
    N = states.shape[0]
    out = torch.empty(N, dtype=torch.complex64, device=model.device)
    if N == 0:
        return out
    lib = _lib.load()
    ws = torch.empty(lib.qmc_energy_workspace_floats(h.ptr, N), dtype=torch.float32, device=model.device)
    _lib.check(h.ptr, lib.qmc_local_energy(
        h.ptr, hamiltonian, float(h_field), states.data_ptr(), N, ws.data_ptr(), out.data_ptr(),
        moments.data_ptr() if moments is not None else None, _stream_ptr(model.device)),
        "qmc_local_energy")
    return out


@scope_op()
def ising_energy(model, states, system_shape=None, K=None, H=None, moments=None):
    """``mcmc_tf.py:59-90``: TFIM local energy PER SPIN, complex64 (N,).
    ``K`` (the receptive field) is implied by the model and only checked."""
    if K is not None and K != model.r:
        raise _lib.QmcError("ising_energy: K=%d differs from the model's receptive field %d" % (K, model.r))
    return _local_energy(model, states, system_shape, _lib.TFIM,
                         globals()["H"] if H is None else H, moments)


@scope_op()
def heisenberg_energy(model, states, system_shape=None, K=None, moments=None):
    """``mcmc_tf.py:93-141``: Marshall-signed AFM Heisenberg local energy per spin."""
    if K is not None and K != model.r:
        raise _lib.QmcError("heisenberg_energy: K=%d differs from the model's receptive field %d"
                            % (K, model.r))
    return _local_energy(model, states, system_shape, _lib.HEISENBERG, 0.0, moments)


@scope_op()
def batched_op(fn, states, batch_size):
    """``mcmc_tf.py:144-153``: apply ``fn`` to chunks of ``batch_size`` rows.  The CUDA
    energy kernel needs no chunking; this keeps the call working (and bounds the
    workspace) for scripts that use it."""
    states = torch.as_tensor(states)
    if states.shape[0] % batch_size:
        raise _lib.QmcError("batched_op: %d rows are not a multiple of batch_size %d"
                            % (states.shape[0], batch_size))
    return torch.cat([fn(states[i:i + batch_size]) for i in range(0, states.shape[0], batch_size)])


@scope_op()
def loss_op(factors, energies):
    """``mcmc_tf.py:35-56``: Re[<E conj(log psi)> - <E><conj(log psi)>] (real scalar)."""
    n = factors.shape[0]
    energies = energies.to(torch.complex64)
    lpc = torch.conj(factors.reshape(n, -1).sum(1))
    e_avg = energies.sum() / n
    return ((energies * lpc).sum() / n - e_avg * lpc.sum() / n).real


def logpsi_gradient(model, states, weights, system_shape=None, out=None):
    """sum_n Re[w_n conj(d log psi_n / d p)] as a flat fp32 vector in ``model.flat`` order.
    With w_n = (E_n - mean E)/N this is d loss_op / d p (``mcmc_tf.py:172-177``)."""
    if _is_sym(model):
        shape = tuple(SYSTEM_SHAPE if system_shape is None else system_shape)
        g = model.gradient(states, weights, shape)
        if out is not None:
            out.add_(g)
            return out
        return g
    if getattr(model, "n_dims", 2) != 2:          # 1-D / 3-D lattices: the generic path
        shape = tuple(SYSTEM_SHAPE if system_shape is None else system_shape)
        d = model.nd_desc(shape)
        st = torch.as_tensor(states, device=model.device).reshape(-1, int(np.prod(shape))).to(torch.int8).contiguous()
        grad = torch.zeros(model.num_params, dtype=torch.float32, device=model.device) if out is None else out
        if st.shape[0]:
            lib = _lib.load()
            wts = weights.to(torch.complex64).contiguous()
            ws = torch.empty(max(lib.qmc_nd_backward_scratch_floats(d, model.device.index or 0, st.shape[0]), 4),
                             dtype=torch.float32, device=model.device)
            _lib.check_nd(lib.qmc_nd_logpsi_backward(d, model.device.index or 0, model.flat.data_ptr(), st.data_ptr(),
                                                     wts.data_ptr(), st.shape[0], ws.data_ptr(), grad.data_ptr(),
                                                     _stream_ptr(model.device)), "qmc_nd_logpsi_backward")
        return grad
    states, system_shape, h = _prep(model, states, system_shape)
    N = states.shape[0]
    grad = torch.zeros(model.num_params, dtype=torch.float32, device=model.device) if out is None else out
    if N == 0:
        return grad
    weights = weights.to(torch.complex64).contiguous()
    lib = _lib.load()
    ws = torch.empty(lib.qmc_backward_workspace_floats(h.ptr, N), dtype=torch.float32, device=model.device)
    _lib.check(h.ptr, lib.qmc_logpsi_backward(h.ptr, states.data_ptr(), weights.data_ptr(), N,
                                              ws.data_ptr(), grad.data_ptr(),
                                              _stream_ptr(model.device)), "qmc_logpsi_backward")
    return grad


class AdamTF1(object):
    """``tf.train.AdamOptimizer`` (``mcmc_tf.py:176``), TF-1 update form:
    lr_t = lr sqrt(1-b2^t)/(1-b1^t);  p -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, flat, lr=LEARNING_RATE, beta1=0.9, beta2=0.999, eps=1e-8):
        self.flat, self.lr, self.b1, self.b2, self.eps = flat, lr, beta1, beta2, eps
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.t = 0

    def step(self, grad):
        self.t += 1
        self.m.mul_(self.b1).add_(grad, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(grad, grad, value=1 - self.b2)
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        self.flat.addcdiv_(self.m, self.v.sqrt().add_(self.eps), value=-lr_t)


def _event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


class EnergiesOp(object):
    """First element of what ``optimize_op`` returns (``mcmc_tf.py:157-179`` returns
    ``(energies, train_op)``): the local energies of the samples the most recent
    ``train_op.run()`` drew.  ``eval()`` / ``numpy()`` fetch them."""

    def __init__(self, step):
        self._step = step

    def eval(self):
        if self._step.last_energies is None:
            raise _lib.QmcError("energies: run the train_op first (sess.run(optimize) runs both together)")
        return self._step.last_energies

    def numpy(self):
        return self.eval().cpu().numpy()


class OptimizeStep(object):
    """What ``optimize_op`` returns.  It unpacks like the reference's return value,

        energies, train_op = optimize_op(sampler, model, energy_fn)        # mcmc_tf.py:157-179

    and ``run()`` is the eager stand-in for ``sess.run(optimize, feed_dict={sampler.new_samples: ...})``
    (``mcmc_tf.py:218-222``): one VMC iteration - sample, local energies, gradient of ``loss_op``,
    TF-1 Adam - returning the sampled local energies.  Under ``torch.distributed`` (one rank per GPU,
    chains sharded) the energy moments and the gradient are all-reduced over NCCL - the only collectives."""

    def __init__(self, sampler, model, energy_fn, learning_rate=LEARNING_RATE, group=None):
        self.sampler, self.model, self.energy_fn = sampler, model, energy_fn
        self.optimizer = AdamTF1(model.flat, learning_rate)
        self.group = group
        self.last_grad = None
        self.last_energy = None
        self.last_energies = None
        self.energies = EnergiesOp(self)
        self.record_events = False      # bench.py: CUDA events at the phase boundaries of run() -> last_events
        self.last_events = None

    # (energies, train_op) like the reference
    def __iter__(self):
        return iter((self.energies, self))

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.energies, self)[i]

    def run(self, new_samples=None):
        from . import distributed as D
        if new_samples is not None:
            self.sampler.new_samples = bool(new_samples)
        ev = []
        mark = (lambda: ev.append(_event())) if self.record_events else (lambda: None)
        mark()
        self.sampler.mcmc_op()
        mark()
        samples = self.sampler.samples_int8()
        energies = self.energy_fn(samples)
        n_tot, e_mean, _, stderr = D.allreduce_energy_moments(energies, self.group)   # collective 1
        mark()
        self.last_energy = (e_mean, stderr)
        weights = D.vmc_weights(energies, e_mean, n_tot)
        grad = logpsi_gradient(self.model, samples, weights, self.sampler.system_shape)
        D.allreduce_gradient(grad, self.group)                                        # collective 2
        self.last_grad = grad
        self.optimizer.step(grad)
        mark()
        self.last_events = ev or None       # [start, after sweep, after energies + allreduce, after gradient + Adam]
        self.last_energies = energies
        return energies


def run(fetches, feed_dict=None):
    """Eager stand-in for ``sess.run`` on what ``optimize_op`` / ``eval_op`` return:
    ``e, _ = run(optimize, feed_dict={"new_samples": it == 0})`` (``mcmc_tf.py:220-222``)."""
    new_samples = None if not feed_dict else feed_dict.get("new_samples")
    if isinstance(fetches, OptimizeStep):
        return fetches.run(new_samples=new_samples), None
    if callable(fetches):
        return fetches()
    raise _lib.QmcError("run: expected the result of optimize_op or a callable")


@scope_op()
def optimize_op(sampler, model, energy_fn, learning_rate=LEARNING_RATE, group=None):
    """``mcmc_tf.py:156-179``: returns ``(energies, train_op)`` - an :class:`OptimizeStep` that unpacks
    into the two; ``train_op.run()`` applies one Adam update and returns the sampled local energies."""
    return OptimizeStep(sampler, model, energy_fn, learning_rate, group)


@scope_op()
def eval_op(sampler, model, energy_fn, batch_size=None):
    """``mcmc_tf.py:182-194``: energies of a fresh ``mcmc_op`` in batches."""
    samples = sampler.mcmc_op()
    bs = ENERGY_BATCH_SIZE if batch_size is None else batch_size
    if samples.shape[0] % bs:
        bs = samples.shape[0]
    return batched_op(energy_fn, samples, bs)
