"""BASELINE config 5 at its stated size: Hamiltonian flow (LeapFrog + momentum-affine layers, reference
example/demo_hamiltonian_flow.jl:27-99,139-147) on a 100-D synthetic logistic-regression posterior, joint target
logp(beta) + sum logN(rho; 0, 1) (:117-124).  The state (dim 200) does not fit one thread: csrc/hmc_warp.cu gives a warp to each
sample.  Checked against the oracle's autograd (which differentiates through the score like the reference's AD does)."""
import math

import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, rel_err, z0

pytestmark = pytest.mark.gpu

TDT = {np.float32: torch.float32, np.float64: torch.float64}


def _flow(tgt, nlayers, nsteps, dtype, seed=0, eps=0.02):
    of = O.hamiltonian_flow(tgt, nlayers, nsteps, math.log(eps), dtype=TDT[dtype])
    rng = np.random.default_rng(seed)
    th = of.theta().double().numpy()
    of.set_theta(torch.from_numpy(th + 0.05 * rng.standard_normal(th.size)).to(TDT[dtype]))
    return of


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("h,n_data,nlayers,nsteps,N", [(100, 256, 3, 3, 48), (100, 300, 15, 3, 8), (33, 50, 4, 2, 70), (128, 64, 2, 1, 33)],
                         ids=["h100-3x3", "h100-demo-depth", "h33", "h128"])
def test_hamiltonian_logreg_large(gpu, h, n_data, nlayers, nsteps, N, dtype):
    nf = gpu
    tgt = O.synthetic_logreg(h, n_data)
    of = _flow(tgt, nlayers, nsteps, dtype)
    jt = O.JointTarget(tgt)
    xs = z0(N, 2 * h, dtype)
    # truth: the float64 oracle at the same (float32-representable) parameters and draws
    of64 = _flow(tgt, nlayers, nsteps, np.float64)
    of64.set_theta(of.theta().double())
    x64 = torch.from_numpy(xs).double()
    v64, g64 = O.elbo_value_and_grad(of64, jt, of64.theta(), x64)
    y64, ld64 = of64.forward(x64)
    gf = gpu_flow(nf, of, dtype)
    y, ld = gf.with_logabsdet_jacobian(xs)
    v, g = nf.api._elbo_impl(gf, gpu_target(nf, jt), xs, want_grad=True)
    terms = nf.batched_elbos(gf, gpu_target(nf, jt), xs)
    tv, tg, ty = (1e-9, 1e-7, 1e-10) if dtype == np.float64 else (1e-5, 1e-4, 2e-5)
    assert rel_err(y, y64.detach().numpy()) <= ty
    assert rel_err(ld, ld64.detach().numpy()) <= max(ty, 1e-6) or np.abs(ld - ld64.detach().numpy()).max() <= 1e-5
    assert abs(v - v64) <= tv * max(abs(v64), 1.0), (v, v64)
    assert rel_err(g, g64) <= tg, rel_err(g, g64)
    assert abs(float(np.mean(terms)) - v64) <= 10 * tv * max(abs(v64), 1.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("tname", ["funnel", "banana", "diag"])
def test_hamiltonian_funnel_large(gpu, tname, dtype):
    """The demo's own target families at a size only the warp-per-sample kernel takes (h = 40, not a power of two)."""
    nf = gpu
    rng = np.random.Generator(np.random.PCG64(3))
    tgt = {"funnel": O.Funnel(40, -2.0, 3.0), "banana": O.Banana(40, 0.3, 4.0),
           "diag": O.DiagNormal(rng.standard_normal(40), rng.uniform(0.5, 1.5, 40))}[tname]
    of = _flow(tgt, 5, 3, dtype, seed=2, eps=0.05)
    jt = O.JointTarget(tgt)
    xs = z0(64, 80, dtype, seed=4)
    of64 = _flow(tgt, 5, 3, np.float64, seed=2, eps=0.05)
    of64.set_theta(of.theta().double())
    v64, g64 = O.elbo_value_and_grad(of64, jt, of64.theta(), torch.from_numpy(xs).double())
    v, g = nf.api._elbo_impl(gpu_flow(nf, of, dtype), gpu_target(nf, jt), xs, want_grad=True)
    tv, tg = (1e-9, 1e-7) if dtype == np.float64 else (1e-5, 1e-4)
    assert abs(v - v64) <= tv * max(abs(v64), 1.0), (v, v64)
    assert rel_err(g, g64) <= tg, rel_err(g, g64)


def test_hamiltonian_100d_device_draws_and_training_step(gpu):
    """z0 = NULL (device Philox draws) through the same kernel, and a short Adam run that must raise the ELBO."""
    nf = gpu
    nf.seed(5)
    ot = O.synthetic_logreg(100, 256)
    tgt = nf.LogReg(ot.X.numpy(), ot.y.numpy(), ot.sigma0)
    of = _flow(ot, 4, 2, np.float64, eps=0.01)
    gf = gpu_flow(nf, of, np.float64)
    jt = gpu_target(nf, O.JointTarget(ot))
    v1, g1 = nf.api._elbo_impl(gf, jt, 512, want_grad=True, seed=7)
    v2, g2 = nf.api._elbo_impl(gf, jt, 512, want_grad=True, seed=7)
    assert v1 == v2 and np.isfinite(v1) and np.all(np.isfinite(g1))
    theta = gf.theta.copy()
    for it in range(30):
        v, g = nf.api._elbo_impl(gf, jt, 256, want_grad=True, seed=100 + it)
        theta = theta + 1e-4 * g / max(np.linalg.norm(g), 1e-12) * np.sqrt(theta.size)
        gf.theta = theta
    v_end, _ = nf.api._elbo_impl(gf, jt, 512, want_grad=True, seed=7)
    assert v_end > v1, (v1, v_end)
    del tgt


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_hamiltonian_100d_inverse_and_logpdf(gpu, dtype):
    """Reversible and volume preserving (reference test/flow.jl:25-39 applied to the demo's bijector), logpdf against the oracle."""
    nf = gpu
    tgt = O.synthetic_logreg(100, 256)
    of = _flow(tgt, 4, 3, dtype)
    of64 = _flow(tgt, 4, 3, np.float64)
    of64.set_theta(of.theta().double())
    gf = gpu_flow(nf, of, dtype)
    xs = z0(40, 200, dtype, seed=3)
    y, lj = gf.with_logabsdet_jacobian(xs)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    rt = 1e-10 if dtype == np.float64 else 2e-4
    np.testing.assert_allclose(xr, xs, rtol=rt, atol=rt)
    np.testing.assert_allclose(lj, -lji, rtol=rt, atol=rt)
    lp = gf.logpdf(y)
    lp64 = of64.logpdf(torch.from_numpy(y).double()).detach().numpy()
    np.testing.assert_allclose(lp, lp64, rtol=1e-9 if dtype == np.float64 else 1e-4, atol=1e-9 if dtype == np.float64 else 1e-3)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("tname", ["funnel", "logreg"])
def test_warp_kernel_matches_thread_kernel(gpu, tname, dtype):
    """Sizes both kernels take (h = 8): the warp-per-sample kernel (forced) against the one-thread-per-sample fused kernel."""
    nf = gpu
    lib = nf._capi.lib()
    tgt = O.Funnel(8, -2.0, 3.0) if tname == "funnel" else O.synthetic_logreg(8, 40)
    of = _flow(tgt, 3, 2, dtype, eps=0.05)
    jt = O.JointTarget(tgt)
    xs = z0(100, 16, dtype)
    res = {}
    try:
        for force in (0, 1):
            nf._capi.check(lib.nf_set_option(b"hmc_warp", force))
            gf = gpu_flow(nf, of, dtype)
            v, g = nf.api._elbo_impl(gf, gpu_target(nf, jt), xs, want_grad=True)
            y, ld = gf.with_logabsdet_jacobian(xs)
            xr, lji = gf.inverse_with_logabsdet_jacobian(y)
            res[force] = (v, g, y, xr)
    finally:
        nf._capi.check(lib.nf_set_option(b"hmc_warp", 0))
    tol = 1e-11 if dtype == np.float64 else 2e-5
    assert abs(res[0][0] - res[1][0]) <= tol * max(abs(res[0][0]), 1.0)
    assert rel_err(res[1][1], res[0][1]) <= (1e-9 if dtype == np.float64 else 1e-4)
    assert rel_err(res[1][2], res[0][2]) <= tol and rel_err(res[1][3], res[0][3]) <= (1e-9 if dtype == np.float64 else 2e-4)


@pytest.mark.parametrize("on_device", [False, True], ids=["host-adam", "device-adam"])
def test_hamiltonian_100d_train_flow(gpu, on_device):
    """`train_flow(elbo, flow, logp_joint, n)` of the demo (example/demo_hamiltonian_flow.jl:160-170) at the 100-D size, host and
    on-device Adam: finite losses, parameters move, ELBO does not get worse on fresh draws."""
    nf = gpu
    ot = O.synthetic_logreg(100, 256)
    tgt = nf.LogReg(ot.X.numpy(), ot.y.numpy(), ot.sigma0)
    flow = nf.hamiltonian_flow(tgt, nlayers=3, L=2, log_eps0=math.log(0.01), paramtype=np.float64)
    jt = nf.JointTarget(tgt)
    v0, _ = nf.api._elbo_impl(flow, jt, 512, want_grad=True, seed=11)
    trained, stats, _ = nf.train_flow(np.random.default_rng(1), nf.elbo, flow, jt, 128, max_iters=40, optimiser=nf.Adam(1e-3),
                                      ADbackend=nf.AutoNFCUDA(on_device=on_device), show_progress=False)
    assert len(stats) == 40 and all(np.isfinite(s["loss"]) for s in stats)
    assert np.abs(trained.theta - flow.theta).max() > 0
    v1, _ = nf.api._elbo_impl(trained, jt, 512, want_grad=True, seed=11)
    assert v1 > v0, (v0, v1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("h,n_data", [(100, 256), (33, 50)], ids=["h100", "h33"])
def test_hamiltonian_large_loglikelihood_grad(gpu, h, n_data, dtype):
    """Forward-KL objective (reference src/objectives/loglikelihood.jl:26-33) and its gradient through the inverse direction."""
    import ctypes as C
    nf = gpu
    tgt = O.synthetic_logreg(h, n_data)
    of = _flow(tgt, 3, 2, dtype)
    of64 = _flow(tgt, 3, 2, np.float64)
    of64.set_theta(of.theta().double())
    gf = gpu_flow(nf, of, dtype)
    rng = np.random.Generator(np.random.PCG64(11))
    xs = (0.7 * rng.standard_normal((60, 2 * h))).astype(dtype)
    v64, g64 = O.loglik_value_and_grad(of64, of64.theta(), torch.from_numpy(xs).double())
    K = nf._capi
    val = C.c_double()
    g = np.empty(gf.theta.size, dtype=dtype)
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), xs.shape[0], K.ptr(xs), 1.0, C.byref(val), K.ptr(g)))
    tv, tg = (1e-9, 1e-7) if dtype == np.float64 else (1e-5, 1e-4)
    assert abs(val.value - v64) <= tv * max(abs(v64), 1.0), (val.value, v64)
    assert rel_err(g, g64) <= tg, rel_err(g, g64)
