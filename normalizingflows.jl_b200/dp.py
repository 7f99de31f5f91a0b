"""Sample-sharded data parallelism for the ELBO / log-likelihood step.

One process per GPU.  theta is replicated, the N_total base draws are split into contiguous blocks, each
rank produces the UN-NORMALISED sums [sum_j d elbo_j/d theta (P numbers) ; sum_j elbo_j] over its block and
exactly one all-reduce of P+1 numbers (NCCL over NVLink on the GPU box, gloo in the CPU tests) followed
by a division by N_total finishes the step.  The reference has no distributed layer; this is the only
collective the path needs.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, every sample owned once."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def finish_sums(sums, n_total: int, scale: float = 1.0):
    """(value, grad) from all-reduced sums: scale/N_total * [grad sums ; value sum]."""
    f = scale / float(n_total)
    return float(sums[-1]) * f, sums[:-1] * f


def allreduce_sums(sums_tensor, group=None):
    """In-place SUM all-reduce of the P+1 vector (a torch tensor on the rank's device)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums_tensor, op=dist.ReduceOp.SUM, group=group)
    return sums_tensor


class DataParallelObjective:
    """value_and_gradient of scale*mean_j elbo_j over a batch sharded across the ranks of torch.distributed.

    `local_sums(lo, hi)` must return the length-(P+1) un-normalised sums for samples [lo, hi) as a torch
    tensor; on the GPU box it is `nf_elbo_sums_dev` writing into a CUDA tensor, in the CPU tests it is the oracle.
    """

    def __init__(self, local_sums, n_total: int, rank: int, world: int, scale: float = -1.0, group=None):
        self.local_sums, self.n_total, self.rank, self.world, self.scale, self.group = local_sums, n_total, rank, world, scale, group

    def __call__(self):
        lo, hi = shard_range(self.n_total, self.rank, self.world)
        sums = self.local_sums(lo, hi)
        allreduce_sums(sums, self.group)
        return finish_sums(sums, self.n_total, self.scale)
