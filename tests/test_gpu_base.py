"""Full-covariance base q0 = MvNormal(mu, Sigma) (reference ext/NormalizingFlowsCUDAExt.jl:43-47,
test/ext/CUDA/cuda.jl:33-44): sampling moments, ELBO / log-likelihood value + gradient and logpdf against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0

pytestmark = pytest.mark.gpu

TOL = {np.float32: (1e-5, 1e-4), np.float64: (1e-9, 1e-7)}


def _dense(of, seed=5):
    rng = np.random.Generator(np.random.PCG64(seed))
    d = of.dim
    A = rng.standard_normal((d, d)) / np.sqrt(d)
    Sigma = A @ A.T + 0.5 * np.eye(d)
    of.base_chol = torch.from_numpy(np.linalg.cholesky(Sigma)).to(of.dtype)
    of.base_mu = torch.from_numpy(rng.standard_normal(d)).to(of.dtype)
    return of, Sigma


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind,dim,tname,kw", [("realnvp", 8, "diag", dict(hdims=[32, 32], nlayers=2)),
                                               ("realnvp", 64, "funnel", dict(hdims=[256, 256], nlayers=1)),
                                               ("nsf", 16, "cross", dict(hdims=[32, 32], K=10, B=5.0, nlayers=2)),
                                               ("planar", 5, "diag", dict(nlayers=6)), ("radial", 2, "warped", dict(nlayers=8))])
def test_dense_base_elbo(gpu, kind, dim, tname, kw, dtype):
    nf = gpu
    of, _ = _dense(oracle_flow(kind, dim, dtype, **kw))
    ot = oracle_target(tname, dim)
    xs = of.base_sample(torch.from_numpy(z0(300, dim, dtype))).numpy().astype(dtype)      # x0 ~ q0
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    gf = gpu_flow(nf, of, dtype)
    v, g = nf.api._elbo_impl(gf, gpu_target(nf, ot), xs, want_grad=True)
    tv, tg = TOL[dtype]
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg
    # device draws: finite, and the value is close to the host-draw value (same distribution, different sample)
    v2 = nf.api._elbo_impl(gf, gpu_target(nf, ot), 4096, seed=3)
    assert np.isfinite(v2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_dense_base_loglikelihood_and_logpdf(gpu, dtype):
    nf = gpu
    of, _ = _dense(oracle_flow("realnvp", 8, dtype, hdims=[32, 32], nlayers=2))
    gf = gpu_flow(nf, of, dtype)
    ys = (0.7 * z0(200, 8, np.float64, seed=3)).astype(dtype)
    v_ref, g_ref = O.loglik_value_and_grad(of, of.theta(), torch.from_numpy(ys))
    K = nf._capi
    val = C.c_double()
    g = np.empty(gf.theta.size, dtype=dtype)
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), 200, K.ptr(ys), 1.0, C.byref(val), K.ptr(g)))
    tv, tg = TOL[dtype]
    assert abs(val.value - v_ref) <= tv * max(abs(v_ref), 1.0)
    assert rel_err(g, g_ref) <= tg
    lp = gf.logpdf(ys)
    np.testing.assert_allclose(lp, of.logpdf(torch.from_numpy(ys)).detach().numpy(), rtol=1e-4, atol=1e-4)


def test_dense_base_sampling_moments(gpu):
    """reference test/ext/CUDA/cuda.jl:33-44: draws of MvNormal(mu, Sigma) on the device have the right mean and covariance."""
    nf = gpu
    of, Sigma = _dense(oracle_flow("realnvp", 8, np.float32, hdims=[32, 32], nlayers=1))
    gf = gpu_flow(nf, of, np.float32)
    zs = gf.rand_base(400000, seed=11).astype(np.float64)
    mu = of.base_mu.double().numpy()
    assert np.abs(zs.mean(0) - mu).max() < 0.02
    assert np.abs(np.cov(zs.T) - Sigma).max() < 0.03
    assert not np.array_equal(gf.rand_base(16, seed=1), gf.rand_base(16, seed=2))


@pytest.mark.parametrize("kind,dim", [("planar", 5), ("radial", 2)])
def test_dense_base_elementwise_inverse_direction(gpu, kind, dim):
    """logpdf / loglikelihood of purely planar / radial flows over a full-covariance base (routed through the layered path)."""
    nf = gpu
    dtype = np.float64
    of, _ = _dense(oracle_flow(kind, dim, dtype, nlayers=5))
    gf = gpu_flow(nf, of, dtype)
    ys = (0.7 * z0(100, dim, np.float64, seed=3)).astype(dtype)
    np.testing.assert_allclose(gf.logpdf(ys), of.logpdf(torch.from_numpy(ys)).detach().numpy(), rtol=1e-7, atol=1e-7)
    v_ref, g_ref = O.loglik_value_and_grad(of, of.theta(), torch.from_numpy(ys))
    K = nf._capi
    val = C.c_double()
    g = np.empty(gf.theta.size, dtype=dtype)
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), 100, K.ptr(ys), 1.0, C.byref(val), K.ptr(g)))
    assert abs(val.value - v_ref) <= 1e-9 * max(abs(v_ref), 1.0) and rel_err(g, g_ref) <= 1e-7
