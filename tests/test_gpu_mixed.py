"""Flows that compose planar / radial layers with coupling layers (`create_flow(Ls, q0) = transformed(q0,
reduce(∘, Ls))`, reference src/flows/utils.jl:23-26, accepts any list of bijectors).  Runs of consecutive elementwise
layers go through the fused elementwise kernel as one segment of the layered sweep."""
import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import TDT, gpu_flow, gpu_target, oracle_target, rel_err, z0

pytestmark = pytest.mark.gpu

TOL = {np.float32: (1e-5, 1e-4), np.float64: (1e-9, 1e-7)}


def mixed_flow(dim, dtype, pattern, hdims=(32, 32), seed=9):
    """pattern: string of P (planar), R (radial), A / B (affine coupling with mask 1:2:d / 2:2:d), S / T (spline couplings),
    H (Shift).  Parameters of planar / radial layers are scaled down so that deep mixes stay well conditioned."""
    rng = np.random.Generator(np.random.PCG64(seed))
    td = TDT[dtype]
    rn = lambda n: torch.from_numpy(0.5 * rng.standard_normal(n)).to(td)   # noqa: E731
    Ls = []
    for ch in pattern:
        if ch == "P":
            Ls.append(O.Planar(rn(dim), rn(dim), rn(1)))
        elif ch == "R":
            Ls.append(O.Radial(rn(1), rn(1), rn(dim)))
        elif ch == "H":
            Ls.append(O.Shift(rn(dim)))
        elif ch in "AB":
            mask = list(range(0 if ch == "A" else 1, dim, 2))
            c = len(mask)
            Ls.append(O.AffineCoupling(dim, mask, O.fnn(rng, dim - c, hdims, c, "tanh", td), O.fnn(rng, dim - c, hdims, c, None, td)))
        elif ch in "ST":
            mask = list(range(0 if ch == "S" else 1, dim, 2))
            c = len(mask)
            Ls.append(O.NeuralSplineCoupling(dim, 6, 4.0, mask, O.fnn(rng, dim - c, hdims, (3 * 6 - 1) * c, None, td)))
    return O.Flow(dim, Ls, dtype=td)


CASES = [(4, "PABR", "diag", (16, 16)), (8, "RRAPBHP", "diag", (32, 32)), (5, "APB", "diag", (32, 32)),
         (6, "PSTR", "diag", (32, 32)), (64, "PABP", "funnel", (256, 256)), (2, "ABPPPRRR", "banana", (16, 16))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("dim,pattern,tname,hd", CASES, ids=[f"d{c[0]}-{c[1]}" for c in CASES])
def test_mixed_flow_elbo_and_forward(gpu, dim, pattern, tname, hd, dtype):
    nf = gpu
    of = mixed_flow(dim, dtype, pattern, hd)
    ot = oracle_target(tname, dim)
    xs = z0(300, dim, dtype)
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    gf = gpu_flow(nf, of, dtype)
    v, g = nf.api._elbo_impl(gf, gpu_target(nf, ot), xs, want_grad=True)
    tv, tg = TOL[dtype]
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)
    y, ld = gf.with_logabsdet_jacobian(xs)
    y_ref, ld_ref = of.forward(torch.from_numpy(xs))
    assert rel_err(y, y_ref.detach().numpy()) <= 10 * tv and rel_err(ld, ld_ref.detach().numpy()) <= 10 * tg
    ys = gf.rand(1000, seed=4)
    assert ys.shape == (1000, dim) and np.all(np.isfinite(ys))


def test_mixed_flow_two_phase_api(gpu):
    """nf_forward_stash / nf_backward with a caller-supplied d/dy and d/dlogdet through a mixed flow."""
    nf = gpu
    dtype = np.float64
    of = mixed_flow(6, dtype, "PABR", (16, 16))
    gf = gpu_flow(nf, of, dtype)
    xs = z0(200, 6, dtype)
    rng = np.random.Generator(np.random.PCG64(1))
    gy, gld = rng.standard_normal((200, 6)), rng.standard_normal(200)
    theta = of.set_theta(of.theta(), requires_grad=True)
    y, ld = of.forward(torch.from_numpy(xs))
    loss = (y * torch.from_numpy(gy)).sum() + (ld * torch.from_numpy(gld)).sum()
    g_ref, = torch.autograd.grad(loss, theta)
    nf.forward_stash(gf, xs)
    g = nf.backward(gf, gy, gld)
    assert rel_err(g, g_ref.numpy()) <= 1e-7


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("dim,pattern,hd", [(4, "PABR", (16, 16)), (8, "RRAPBHP", (32, 32)), (6, "PSTR", (32, 32)), (64, "RABP", (256, 256))],
                         ids=["d4-PABR", "d8-RRAPBHP", "d6-PSTR", "d64-RABP"])
def test_mixed_flow_inverse_direction(gpu, dim, pattern, hd, dtype):
    """loglikelihood value + gradient, logpdf and the inverse round trip (reference src/objectives/loglikelihood.jl:26-33,
    test/flow.jl:25-39) of mixed flows: the inverse sweep runs the same segments through the inverse elementwise kernel."""
    import ctypes as C
    nf = gpu
    of = mixed_flow(dim, dtype, pattern, hd)
    gf = gpu_flow(nf, of, dtype)
    ys = (0.7 * z0(200, dim, np.float64, seed=3)).astype(dtype)
    v_ref, g_ref = O.loglik_value_and_grad(of, of.theta(), torch.from_numpy(ys))
    K = nf._capi
    val = C.c_double()
    g = np.empty(gf.theta.size, dtype=dtype)
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), 200, K.ptr(ys), 1.0, C.byref(val), K.ptr(g)))
    tv, tg = (2e-5, 5e-4) if dtype == np.float32 else (1e-9, 1e-7)     # Float32: chained scalar root finds (planar inverse)
    assert abs(val.value - v_ref) <= tv * max(abs(v_ref), 1.0), (val.value, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)
    np.testing.assert_allclose(gf.logpdf(ys), of.logpdf(torch.from_numpy(ys)).detach().numpy(), rtol=1e-4, atol=1e-4)
    xs = z0(50, dim, dtype, seed=8)
    y, lj = gf.with_logabsdet_jacobian(xs)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    rt = 5e-4 if dtype == np.float32 else 1e-8
    np.testing.assert_allclose(xr, xs, rtol=rt, atol=rt)
    np.testing.assert_allclose(lj, -lji, rtol=rt, atol=rt)
