"""world_size-2 gloo test of the sample-sharded step: per-rank un-normalised sums +
ONE all-reduce of P+1 numbers + division by N_total == the single-process value and gradient."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import nfload
    import nf_oracle as O
    from helpers import oracle_flow, oracle_target
    nf = nfload.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    of = oracle_flow("realnvp", 5, np.float64, hdims=[16, 16], nlayers=2)
    ot = oracle_target("diag", 5)
    xs = torch.from_numpy(O.synthetic_z0(n_total, 5)).double()
    theta = of.theta()

    def local_sums(lo, hi):      # stands in for nf_elbo_sums_dev on the rank's shard
        n = hi - lo
        if n == 0:
            return torch.zeros(theta.numel() + 1, dtype=torch.float64)
        v, g = O.elbo_value_and_grad(of, ot, theta, xs[lo:hi])
        return torch.cat([torch.from_numpy(g) * n, torch.tensor([v * n], dtype=torch.float64)])

    step = nf.dp.DataParallelObjective(local_sums, n_total, rank, world, scale=-1.0)
    loss, grad = step()
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([grad.numpy(), [loss]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [101, 3])
def test_sharded_step_matches_single_process(tmp_path, n_total):
    import nf_oracle as O
    from helpers import oracle_flow, oracle_target
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    of = oracle_flow("realnvp", 5, np.float64, hdims=[16, 16], nlayers=2)
    ot = oracle_target("diag", 5)
    xs = torch.from_numpy(O.synthetic_z0(n_total, 5)).double()
    v, g = O.elbo_value_and_grad(of, ot, of.theta(), xs)
    ref = np.concatenate([-g, [-v]])
    r0, r1 = (np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(2))
    assert np.array_equal(r0, r1)                      # every rank ends with the same numbers
    assert np.allclose(r0, ref, rtol=1e-10, atol=1e-12)
