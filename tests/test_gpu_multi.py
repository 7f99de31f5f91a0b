"""The data-parallel entry points behind the C ABI: `nf_elbo_sums_dev` (per-shard un-normalised sums) and
`nf_*_value_and_grad_multi` (shards + the NCCL all-reduce inside libnfcuda).

One-GPU boxes run the single-rank paths and the shard arithmetic; the two-device tests run when the box has >= 2 GPUs
(`gpurun --gpus 2`) and compare the all-reduced gradient with the single-device one."""
import ctypes as C

import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0

pytestmark = pytest.mark.gpu


def _realnvp(nf, dim=8, hd=(32, 32), nlayers=2):
    of = oracle_flow("realnvp", dim, np.float32, hdims=list(hd), nlayers=nlayers)
    return of, gpu_flow(nf, of, np.float32)


def test_elbo_sums_over_halves_equal_the_whole(gpu):
    """sums(first half) + sums(second half) == sums(whole) == N x (value, gradient): the identity the all-reduce relies on."""
    nf = gpu
    K, lib = nf._capi, nf._capi.lib()
    of, gf = _realnvp(nf)
    gt = gpu_target(nf, oracle_target("diag", 8))
    N = 1000
    xs = z0(N, 8, np.float32)
    dev = torch.device("cuda", 0)
    theta = torch.from_numpy(gf.theta).to(dev)
    zs = torch.from_numpy(xs).to(dev)
    P = gf.num_params
    out = [torch.zeros(P + 1, device=dev, dtype=torch.float32) for _ in range(3)]
    torch.cuda.synchronize()
    h, th = gf.handle(), gt.handle()
    K.check(lib.nf_elbo_sums_dev(h, th, theta.data_ptr(), N, zs.data_ptr(), 0, out[0].data_ptr()))
    lo = 437                                             # ragged split
    K.check(lib.nf_elbo_sums_dev(h, th, theta.data_ptr(), lo, zs.data_ptr(), 0, out[1].data_ptr()))
    K.check(lib.nf_elbo_sums_dev(h, th, theta.data_ptr(), N - lo, zs[lo:].data_ptr(), 0, out[2].data_ptr()))
    whole, parts = out[0].cpu().numpy().astype(np.float64), (out[1] + out[2]).cpu().numpy().astype(np.float64)
    assert rel_err(parts, whole) <= 2e-6
    v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    assert abs(whole[P] / N - v) <= 2e-6 * max(abs(v), 1.0)
    assert rel_err(whole[:P] / N, g) <= 2e-6
    # and against the oracle, so the identity is not between two wrong numbers
    v_ref, g_ref = O.elbo_value_and_grad(of, oracle_target("diag", 8), of.theta(), torch.from_numpy(xs))
    assert abs(whole[P] / N - v_ref) <= 1e-5 * max(abs(v_ref), 1.0) and rel_err(whole[:P] / N, g_ref) <= 1e-4


def test_multi_entry_point_with_one_rank_matches_the_plain_call(gpu):
    """nf_elbo_value_and_grad_multi over a one-device communicator == nf_elbo_value_and_grad, for host Z0 and for device
    Philox draws; same for the log-likelihood twin."""
    nf = gpu
    of, gf = _realnvp(nf)
    gt = gpu_target(nf, oracle_target("diag", 8))
    comm = nf.dp.Comm.init_all([0])
    assert comm.size == 1 and comm.local_devices == [0] and comm.local_ranks == [0]
    xs = z0(777, 8, np.float32)
    v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    vm, gm = nf.dp.elbo_value_and_grad_multi(comm, [gf], [gt], gf.theta, 777, xs)
    assert vm == pytest.approx(v, rel=1e-6) and rel_err(gm, g) <= 1e-6
    v, g = nf.api._elbo_impl(gf, gt, 4096, want_grad=True, seed=99)
    vm, gm = nf.dp.elbo_value_and_grad_multi(comm, [gf], [gt], gf.theta, 4096, None, seed=99)
    assert vm == pytest.approx(v, rel=1e-6) and rel_err(gm, g) <= 1e-6
    ys = (0.7 * z0(500, 8, np.float64, seed=3)).astype(np.float32)
    val = C.c_double()
    g1 = np.empty(gf.theta.size, np.float32)
    K = nf._capi
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), 500, K.ptr(ys), 1.0, C.byref(val), K.ptr(g1)))
    vm, gm = nf.dp.loglik_value_and_grad_multi(comm, [gf], gf.theta, 500, ys)
    assert vm == pytest.approx(val.value, rel=1e-6) and rel_err(gm, g1) <= 1e-6
    comm.close()


def test_shard_range_partitions_the_batch(gpu):
    lib = gpu._capi.lib()
    for n, r in ((101, 2), (3, 2), (1 << 20, 8), (1000, 7)):
        prev = 0
        for k in range(r):
            b, e = C.c_int64(), C.c_int64()
            lib.nf_shard_range(n, r, k, C.byref(b), C.byref(e))
            assert b.value == prev and e.value >= b.value
            assert (b.value, e.value) == gpu.dp.shard_range(n, k, r)
            prev = e.value
        assert prev == n


def _two_gpus(nf):
    n = C.c_int()
    nf._capi.check(nf._capi.lib().nf_device_count(C.byref(n)))
    return n.value >= 2


@pytest.mark.parametrize("kind", ["realnvp", "planar"])
def test_two_devices_match_one(gpu, kind):
    """One process drives two GPUs (`nf_comm_init_all`): the all-reduced value and gradient equal the single-device ones
    -- for host-supplied Z0 (sharded rows) and for device Philox draws (each rank draws its rows of the one global matrix)."""
    nf = gpu
    if not _two_gpus(nf):
        pytest.skip("needs two GPUs")
    if kind == "realnvp":
        of = oracle_flow("realnvp", 8, np.float32, hdims=[32, 32], nlayers=2)
        tname, dim = "diag", 8
    else:
        of = oracle_flow("planar", 2, np.float32, nlayers=10)
        tname, dim = "banana", 2
    ot = oracle_target(tname, dim)
    comm = nf.dp.Comm.init_all([0, 1])
    flows, targets = nf.dp.replicate(comm, lambda: gpu_flow(nf, of, np.float32), lambda: gpu_target(nf, ot))
    nf._capi.check(nf._capi.lib().nf_init(0))
    N = 1001                                               # odd: shards of 501 and 500
    xs = z0(N, dim, np.float32)
    v1, g1 = nf.api._elbo_impl(flows[0], targets[0], xs, want_grad=True)
    v2, g2 = nf.dp.elbo_value_and_grad_multi(comm, flows, targets, flows[0].theta, N, xs)
    assert v2 == pytest.approx(v1, rel=2e-6) and rel_err(g2, g1) <= 5e-6
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    assert abs(v2 - v_ref) <= 1e-5 * max(abs(v_ref), 1.0) and rel_err(g2, g_ref) <= 1e-4
    v1, g1 = nf.api._elbo_impl(flows[0], targets[0], 5000, want_grad=True, seed=7)
    v2, g2 = nf.dp.elbo_value_and_grad_multi(comm, flows, targets, flows[0].theta, 5000, None, seed=7)
    assert v2 == pytest.approx(v1, rel=2e-6) and rel_err(g2, g1) <= 5e-6
    comm.close()


def test_two_devices_hamiltonian_100d(gpu):
    """BASELINE config 5 as stated: the 100-D Hamiltonian flow sample-sharded over the devices of one communicator."""
    import math
    nf = gpu
    if not _two_gpus(nf):
        pytest.skip("needs two GPUs")
    tgt = O.synthetic_logreg(100, 256)
    of = O.hamiltonian_flow(tgt, 3, 2, math.log(0.02), dtype=torch.float64)
    jt = O.JointTarget(tgt)
    comm = nf.dp.Comm.init_all([0, 1])
    flows, targets = nf.dp.replicate(comm, lambda: gpu_flow(nf, of, np.float64), lambda: gpu_target(nf, jt))
    nf._capi.check(nf._capi.lib().nf_init(0))
    N = 301
    xs = z0(N, 200, np.float64)
    v1, g1 = nf.api._elbo_impl(flows[0], targets[0], xs, want_grad=True)
    v2, g2 = nf.dp.elbo_value_and_grad_multi(comm, flows, targets, flows[0].theta, N, xs)
    assert v2 == pytest.approx(v1, rel=1e-12) and rel_err(g2, g1) <= 1e-11
    v_ref, g_ref = O.elbo_value_and_grad(of, jt, of.theta(), torch.from_numpy(xs))
    assert abs(v2 - v_ref) <= 1e-9 * max(abs(v_ref), 1.0) and rel_err(g2, g_ref) <= 1e-7
    v1, g1 = nf.api._elbo_impl(flows[0], targets[0], 1000, want_grad=True, seed=7)
    v2, g2 = nf.dp.elbo_value_and_grad_multi(comm, flows, targets, flows[0].theta, 1000, None, seed=7)
    assert v2 == pytest.approx(v1, rel=1e-12) and rel_err(g2, g1) <= 1e-11
    comm.close()
