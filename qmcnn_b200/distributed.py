"""Data parallelism over chains: the only cross-rank steps of the VMC hot path.

Chains are independent given the parameters (``sampler.py:117-133`` has no cross-row
term), energies are per sample (``mcmc_tf.py:86-90``) and the loss is a mean over samples
(``mcmc_tf.py:53-55``), so rank g of G owns global chains [g*S, (g+1)*S), parameters are
replicated, and one iteration needs exactly two all-reduces:
the energy moments and the gradient.  The reference has no distributed code; this is new.

Everything here is device-agnostic torch (NCCL on GPUs, gloo in the CPU tests).
"""
import torch


def world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def chain_partition(num_chains_total, rank, world_size):
    """Global chain ids [first, first + count) owned by ``rank``: contiguous blocks, remainder
    spread over the first ranks.  The in-kernel Philox stream is keyed by the GLOBAL chain id,
    so results do not depend on how many ranks the chains are split over."""
    base, rem = divmod(int(num_chains_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_energy_moments(energies, group=None):
    """[N, sum Re E, sum Im E, sum |E|^2] over all ranks (float64).  Returns
    (n_total, mean (complex128 scalar tensor), variance of Re E, stderr of the mean)."""
    import torch.distributed as dist
    e = energies.to(torch.complex128)
    mom = torch.stack([torch.tensor(float(e.numel()), dtype=torch.float64, device=e.device),
                       e.real.sum(), e.imag.sum(), (e.real ** 2).sum()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mom, group=group)
    n = mom[0]
    mean = torch.complex(mom[1] / n, mom[2] / n)
    var = torch.clamp(mom[3] / n - (mom[1] / n) ** 2, min=0.0)
    return n, mean, var, torch.sqrt(var / n)


def vmc_weights(energies, mean, n_total):
    """w_n = (E_n - <E>) / N with the GLOBAL mean and count: summing the per-rank
    gradients of these weights gives d loss_op / d p of the full batch."""
    return ((energies.to(torch.complex128) - mean) / n_total).to(torch.complex64)


def allreduce_gradient(grad, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad, group=group)
    return grad
