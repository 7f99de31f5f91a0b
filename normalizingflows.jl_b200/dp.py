"""Sample-sharded data parallelism for the ELBO / log-likelihood step.

theta is replicated, the N_total base draws are split into contiguous blocks, each rank produces the UN-NORMALISED sums
[sum_j d elbo_j/d theta (P numbers) ; sum_j elbo_j] over its block and exactly one all-reduce of P+1 numbers followed by a
division by N_total finishes the step.  The reference has no distributed layer; this is the only collective the path needs.

On the GPU the whole step -- shards, the NCCL all-reduce on each device's compute stream, the scaling -- is ONE call into
libnfcuda (`Comm` + `elbo_value_and_grad_multi` below; include/nfcuda.h `nf_*_multi`): a host in any language drives G
devices from one process (`Comm.init_all`) or joins as one rank per process (`Comm.init_rank`; the launcher only carries
the 128-byte NCCL id).  `DataParallelObjective` keeps the same arithmetic over `torch.distributed` for the CPU (gloo)
tests of the host logic, with the oracle standing in for the device sums.
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


# ---------------------------------------------------------------------------------------------
# the GPU path: communicator + one-call data-parallel objective behind the C ABI
# ---------------------------------------------------------------------------------------------
class Comm:
    """nf_comm_t: the ranks of a data-parallel job and the local devices this process drives."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def init_all(cls, devices):
        """One process, `devices` GPUs (`ncclCommInitAll`) -- what a Julia host does."""
        import ctypes as C
        from . import _capi as K
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        K.check(K.lib().nf_comm_init_all(C.byref(h), len(devices), devs))
        return cls(h)

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _capi as K
        buf = C.create_string_buffer(K.NF_UNIQUE_ID_BYTES)
        K.check(K.lib().nf_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def init_rank(cls, n_ranks: int, rank: int, uid: bytes, device: int):
        """One process per GPU: every rank passes the id rank 0 created (`ncclCommInitRank`)."""
        import ctypes as C
        from . import _capi as K
        h = C.c_void_p()
        buf = C.create_string_buffer(uid, K.NF_UNIQUE_ID_BYTES) if uid is not None else None
        K.check(K.lib().nf_comm_init_rank(C.byref(h), n_ranks, rank, buf, device))
        return cls(h)

    @property
    def size(self):
        from . import _capi as K
        return K.lib().nf_comm_size(self._h)

    @property
    def local_devices(self):
        from . import _capi as K
        return [K.lib().nf_comm_local_device(self._h, i) for i in range(K.lib().nf_comm_local_size(self._h))]

    @property
    def local_ranks(self):
        from . import _capi as K
        return [K.lib().nf_comm_local_rank(self._h, i) for i in range(K.lib().nf_comm_local_size(self._h))]

    def close(self):
        from . import _capi as K
        if self._h:
            K.lib().nf_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def replicate(comm: "Comm", make_flow, make_target=None):
    """One (flow, target) replica per local device of `comm`: `make_flow()` / `make_target()` are called with that device
    current (handles are created lazily, so they are forced here)."""
    from . import _capi as K
    flows, targets = [], []
    for dev in comm.local_devices:
        K.check(K.lib().nf_init(dev))
        f = make_flow()
        f.handle()
        flows.append(f)
        if make_target is not None:
            t = make_target()
            t.handle()
            targets.append(t)
    return flows, targets


def _handles(objs):
    import ctypes as C
    return (C.c_void_p * len(objs))(*[o.handle() for o in objs])


def elbo_value_and_grad_multi(comm: "Comm", flows, targets, theta, n_total: int, z0=None, seed: int = 0, scale: float = 1.0):
    """(value, grad) of scale * mean ELBO over n_total samples sharded across every device of the job (nf_elbo_value_and_grad_multi).
    z0: this process's rows ([n_local, dim], all rows with `init_all`) or None for device Philox draws."""
    import ctypes as C
    from . import _capi as K
    theta = np.ascontiguousarray(theta, dtype=flows[0].paramtype)
    grad = np.empty(theta.size, dtype=flows[0].paramtype)
    val = C.c_double()
    zs = None if z0 is None else np.ascontiguousarray(z0, dtype=flows[0].paramtype)
    K.check(K.lib().nf_elbo_value_and_grad_multi(comm._h, _handles(flows), _handles(targets), K.ptr(theta), int(n_total), K.ptr(zs),
                                                 int(seed), float(scale), C.byref(val), K.ptr(grad)))
    return val.value, grad


def loglik_value_and_grad_multi(comm: "Comm", flows, theta, n_total: int, xs, scale: float = 1.0):
    """Forward-KL twin (nf_loglik_value_and_grad_multi): xs = this process's rows of the data batch."""
    import ctypes as C
    from . import _capi as K
    theta = np.ascontiguousarray(theta, dtype=flows[0].paramtype)
    grad = np.empty(theta.size, dtype=flows[0].paramtype)
    val = C.c_double()
    xs = np.ascontiguousarray(xs, dtype=flows[0].paramtype)
    K.check(K.lib().nf_loglik_value_and_grad_multi(comm._h, _handles(flows), K.ptr(theta), int(n_total), K.ptr(xs), float(scale),
                                                   C.byref(val), K.ptr(grad)))
    return val.value, grad
