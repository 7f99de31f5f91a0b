"""Unusual shapes through the C ABI vs the oracle: hidden widths that are not multiples of the 64-column plane tiles, three-layer
and one-layer conditioners, widths above 256 (beyond the tcgen05 wgrad tile), arbitrary (non-alternating) coupling masks,
spline bin counts across the KMAX template buckets, odd dimensions."""
import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_target, rel_err, z0

pytestmark = pytest.mark.gpu
TDT = {np.float32: torch.float32, np.float64: torch.float64}


def _affine_flow(dim, hdims, masks, dtype, seed=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    layers = []
    for m in masks:
        c = len(m)
        layers.append(O.AffineCoupling(dim, list(m), O.fnn(rng, dim - c, hdims, c, "tanh", TDT[dtype]),
                                       O.fnn(rng, dim - c, hdims, c, None, TDT[dtype])))
    return O.Flow(dim, layers, dtype=TDT[dtype])


def _spline_flow(dim, hdims, K, B, masks, dtype, seed=2):
    rng = np.random.Generator(np.random.PCG64(seed))
    layers = [O.NeuralSplineCoupling(dim, K, B, list(m), O.fnn(rng, dim - len(m), hdims, (3 * K - 1) * len(m), None, TDT[dtype]))
              for m in masks]
    return O.Flow(dim, layers, dtype=TDT[dtype])


def _check(nf, of, of64, ot, N, dtype, tv, tg):
    xs = z0(N, of.dim, dtype, seed=N + of.dim)
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    if dtype == np.float32:   # Float32 noise floor of the CPU path itself
        th = of.theta().clone()
        v64, g64 = O.elbo_value_and_grad(of64, ot, th.double(), torch.from_numpy(xs).double())
        tv, tg = max(tv, 2 * abs(v_ref - v64) / max(abs(v64), 1.0)), max(tg, 2 * rel_err(g_ref, g64))
    v, g = nf.api._elbo_impl(gpu_flow(nf, of, dtype), gpu_target(nf, ot), xs, want_grad=True)
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)


AFFINE = [
    (7, [17], [(0, 3, 4), (1, 2, 5, 6)]),                 # one hidden layer, width 17, irregular masks
    (6, [48, 100, 33], [(0, 1, 2), (3, 4, 5)]),           # three hidden layers, block masks
    (9, [200, 64], [(8,), (0, 1, 2, 3, 4, 5, 6, 7)]),     # a single transformed coordinate / a single conditioner input
    (12, [300, 40], [tuple(range(0, 12, 2)), tuple(range(1, 12, 2))]),   # width 300 > 256
    (70, [96, 96], [tuple(range(0, 70, 2)), tuple(range(1, 70, 2))]),    # 35 conditioner inputs (K padded 35 -> 64), d > 64
    (8, [512, 300], [(0, 1, 2, 3), (4, 5, 6, 7)]),        # 512 x 300 Dense: 2 x 2 blocks of the tcgen05 weight-gradient tile
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("dim,hdims,masks", AFFINE, ids=["w17", "w48-100-33", "c1", "w300", "d70", "w512x300"])
def test_affine_coupling_shapes(gpu, dim, hdims, masks, dtype):
    of = _affine_flow(dim, hdims, masks, dtype)
    of64 = _affine_flow(dim, hdims, masks, np.float64)
    tv, tg = (1e-5, 1e-4) if dtype == np.float32 else (1e-9, 1e-7)
    _check(gpu, of, of64, oracle_target("diag", dim), 333, dtype, tv, tg)


SPLINE = [
    (5, [24], 2, 3.0, [(0, 2, 4), (1, 3)]),          # K = 2 (smallest), one hidden layer
    (6, [40, 40], 5, 4.0, [(0, 1, 2), (3, 4, 5)]),
    (4, [32, 32], 16, 5.0, [(0, 3), (1, 2)]),        # KMAX = 16 bucket
    (4, [32, 32], 33, 6.0, [(1,), (0, 2, 3)]),       # KMAX = 64 bucket, 98 logits per coordinate
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("dim,hdims,K,B,masks", SPLINE, ids=["K2", "K5", "K16", "K33"])
def test_spline_coupling_shapes(gpu, dim, hdims, K, B, masks, dtype):
    of = _spline_flow(dim, hdims, K, B, masks, dtype)
    of64 = _spline_flow(dim, hdims, K, B, masks, np.float64)
    tv, tg = (2e-5, 2e-4) if dtype == np.float32 else (1e-9, 1e-7)
    _check(gpu, of, of64, oracle_target("diag", dim), 257, dtype, tv, tg)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("which", ["affine-w17", "affine-w512x300", "spline-K5", "spline-K33"])
def test_loglikelihood_unusual_shapes(gpu, which, dtype):
    """forward-KL objective (inverse sweep + implicit-differentiation backward) on the same unusual shapes."""
    import ctypes as C
    nf = gpu
    if which.startswith("affine"):
        dim, hdims, masks = AFFINE[0] if which == "affine-w17" else AFFINE[5]
        of = _affine_flow(dim, hdims, masks, dtype)
    else:
        dim, hdims, K_, B, masks = SPLINE[1] if which == "spline-K5" else SPLINE[3]
        of = _spline_flow(dim, hdims, K_, B, masks, dtype)
    gf = gpu_flow(nf, of, dtype)
    rng = np.random.Generator(np.random.PCG64(11))
    xs = (0.7 * rng.standard_normal((150, dim))).astype(dtype)
    v_ref, g_ref = O.loglik_value_and_grad(of, of.theta(), torch.from_numpy(xs))
    K = nf._capi
    val = C.c_double()
    g = np.empty(gf.theta.size, dtype=dtype)
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), xs.shape[0], K.ptr(xs), 1.0, C.byref(val), K.ptr(g)))
    tv, tg = (2e-5, 3e-4) if dtype == np.float32 else (1e-9, 1e-7)
    assert abs(val.value - v_ref) <= tv * max(abs(v_ref), 1.0), (val.value, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)
    y, lj = gf.with_logabsdet_jacobian(xs)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    rt = 2e-4 if dtype == np.float32 else 1e-9
    np.testing.assert_allclose(xr, xs, rtol=rt, atol=rt)
    np.testing.assert_allclose(lj, -lji, rtol=rt, atol=rt)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["affine", "spline", "planar"])
def test_non_standard_base_distribution(gpu, kind, dtype):
    """q0 = MvNormal(mu, Diagonal(sigma^2)) with mu != 0, sigma != 1 (reference example/demo_planar_flow.jl:24 style) under
    every path: host-supplied draws from q0, log q0 in the ELBO / log-likelihood heads."""
    nf = gpu
    dim = 6
    rng = np.random.Generator(np.random.PCG64(21))
    mu, sg = rng.standard_normal(dim), rng.uniform(0.5, 2.0, dim)
    if kind == "affine":
        of = _affine_flow(dim, [24, 24], [(0, 2, 4), (1, 3, 5)], dtype)
    elif kind == "spline":
        of = _spline_flow(dim, [24], 6, 4.0, [(0, 1, 2), (3, 4, 5)], dtype)
    else:
        of = O.planarflow(dim, 5, TDT[dtype], rng)
    of.base_mu, of.base_sigma = torch.from_numpy(mu).to(TDT[dtype]), torch.from_numpy(sg).to(TDT[dtype])
    ot = oracle_target("diag", dim)
    xs = (mu + sg * z0(300, dim, np.float64, seed=3)).astype(dtype)          # draws from q0
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    gf = gpu_flow(nf, of, dtype)
    v, g = nf.api._elbo_impl(gf, gpu_target(nf, ot), xs, want_grad=True)
    tv, tg = (2e-5, 2e-4) if dtype == np.float32 else (1e-9, 1e-7)
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)
    ys = (0.8 * z0(200, dim, np.float64, seed=4)).astype(dtype)
    lp = gf.logpdf(ys)
    lp_ref = of.logpdf(torch.from_numpy(ys)).detach().numpy()
    np.testing.assert_allclose(lp, lp_ref, rtol=2e-4 if dtype == np.float32 else 1e-9, atol=2e-4 if dtype == np.float32 else 1e-9)


def test_wide_state_fallback_paths(gpu):
    """d = 300 > 256: the generic coupling kernel and the untiled ELBO head (their row-oriented / shared-memory fast paths stop
    at 256 columns)."""
    dim = 300
    masks = [tuple(range(0, dim, 2)), tuple(range(1, dim, 2))]
    for dtype, tv, tg in ((np.float64, 1e-9, 1e-7), (np.float32, 2e-5, 2e-4)):
        of = _affine_flow(dim, [64], masks, dtype)
        of64 = _affine_flow(dim, [64], masks, np.float64)
        _check(gpu, of, of64, oracle_target("diag", dim), 200, dtype, tv, tg)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["affine", "spline"])
def test_shift_scale_mixed_with_couplings(gpu, kind, dtype):
    """Trainable Shift / Scale layers composed with coupling layers (a `Shift ∘ Scale` pre-conditioner in front of the couplings,
    one more Scale behind them): ELBO and log-likelihood value + gradient, forward / inverse round trip."""
    import ctypes as C
    nf = gpu
    dim = 6
    rng = np.random.Generator(np.random.PCG64(31))
    td = TDT[dtype]
    core = (_affine_flow(dim, [24, 24], [(0, 2, 4), (1, 3, 5)], dtype) if kind == "affine"
            else _spline_flow(dim, [24], 6, 4.0, [(0, 1, 2), (3, 4, 5)], dtype)).layers
    layers = ([O.Scale(torch.from_numpy(rng.uniform(0.6, 1.6, dim)).to(td))] + core +
              [O.Shift(torch.from_numpy(0.3 * rng.standard_normal(dim)).to(td)),
               O.Scale(torch.from_numpy(rng.uniform(0.7, 1.4, dim) * rng.choice([-1.0, 1.0], dim)).to(td))])
    of = O.Flow(dim, layers, dtype=td)
    ot = oracle_target("diag", dim)
    xs = z0(300, dim, dtype, seed=5)
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    gf = gpu_flow(nf, of, dtype)
    v, g = nf.api._elbo_impl(gf, gpu_target(nf, ot), xs, want_grad=True)
    tv, tg = (2e-5, 2e-4) if dtype == np.float32 else (1e-9, 1e-7)
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)
    ys = (0.7 * z0(200, dim, np.float64, seed=6)).astype(dtype)
    vl_ref, gl_ref = O.loglik_value_and_grad(of, of.theta(), torch.from_numpy(ys))
    K = nf._capi
    val = C.c_double()
    gl = np.empty(gf.theta.size, dtype=dtype)
    K.check(K.lib().nf_loglik_value_and_grad(gf.handle(), K.ptr(gf.theta), ys.shape[0], K.ptr(ys), 1.0, C.byref(val), K.ptr(gl)))
    assert abs(val.value - vl_ref) <= tv * max(abs(vl_ref), 1.0), (val.value, vl_ref)
    assert rel_err(gl, gl_ref) <= (3e-4 if dtype == np.float32 else tg), rel_err(gl, gl_ref)
    y, lj = gf.with_logabsdet_jacobian(xs)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    rt = 2e-4 if dtype == np.float32 else 1e-9
    np.testing.assert_allclose(xr, xs, rtol=rt, atol=rt)
    np.testing.assert_allclose(lj, -lji, rtol=rt, atol=rt)
