"""Fused AffineCoupling conditioner kernels (csrc/fused_coupling.cuh: two-team streaming kernel, the default;
csrc/fused_coupling_w128.cuh: 128-column-MMA kernel) against the layer-by-layer tcgen05 path and the float64 oracle.

Reference: src/flows/realnvp.jl:57-110 (coupling forward / inverse), src/flows/utils.jl:71-100 (conditioner MLP).
The fused kernels serve Float32 RealNVP layers with d % 4 == 0, d <= 64 and two equal hidden widths <= 256; every case below
is inside that envelope (everything else keeps the layered path, covered by test_gpu_parity / test_gpu_shapes)."""
import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0

pytestmark = pytest.mark.gpu

VARIANTS = {"two_team": 0, "wide": 1}


def _set(nf, fused, variant=0):
    lib = nf._capi.lib()
    nf._capi.check(lib.nf_set_option(b"fused_coupling", int(fused)))
    nf._capi.check(lib.nf_set_option(b"fused_variant", int(variant)))


@pytest.fixture
def restore(gpu):
    yield
    _set(gpu, 1, 0)


def _pair(dim, hd, nlayers):
    of32 = oracle_flow("realnvp", dim, np.float32, hdims=hd, nlayers=nlayers)
    of64 = oracle_flow("realnvp", dim, np.float64, hdims=hd, nlayers=nlayers)
    of64.set_theta(of32.theta().double())
    return of32, of64


# hidden widths 32 .. 256 cover 1 .. 4 hidden chunks (odd chunk counts exercise the short last accumulation chain); batch sizes
# cover a single row, one short tile, tile + 1 row, and more tiles than one CTA wave would take at small grids
CASES = [(8, [32, 32], 1, 1), (8, [32, 32], 1, 100), (16, [64, 64], 2, 129), (32, [128, 128], 1, 257), (64, [192, 192], 1, 130),
         (64, [256, 256], 2, 300), (64, [256, 256], 4, 1000), (24, [96, 96], 2, 500), (64, [200, 200], 1, 640)]


@pytest.mark.parametrize("variant", list(VARIANTS), ids=list(VARIANTS))
@pytest.mark.parametrize("dim,hd,nlayers,N", CASES, ids=["d%d_h%d_L%d_N%d" % (c[0], c[1][0], c[2], c[3]) for c in CASES])
def test_fused_forward_elbo_and_gradient(gpu, restore, variant, dim, hd, nlayers, N):
    nf = gpu
    of32, of64 = _pair(dim, hd, nlayers)
    ot = oracle_target("funnel", dim)
    gt = gpu_target(nf, ot)
    xs = z0(N, dim, np.float32)
    x64 = torch.from_numpy(xs).double()
    v64, g64 = O.elbo_value_and_grad(of64, ot, of64.theta(), x64)
    y64, ld64 = of64.forward(x64)
    y64, ld64 = y64.detach().numpy(), ld64.detach().numpy()
    out = {}
    for fused in (0, 1):
        _set(nf, fused, VARIANTS[variant])
        gf = gpu_flow(nf, of32, np.float32)
        y, ld = gf.with_logabsdet_jacobian(xs)
        v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
        out[fused] = (y, ld, v, g)
        # north-star tolerances against the float64 oracle (value 1e-5, gradient 1e-4), forward map well inside them
        assert rel_err(y, y64) <= 2e-6 and rel_err(ld, ld64) <= 5e-6
        assert abs(v - v64) <= 1e-5 * max(abs(v64), 1.0)
        assert rel_err(g, g64) <= 1e-4
    # the fused kernel and the layer-by-layer path run the same arithmetic up to accumulation-chain length and rounding order
    assert rel_err(out[1][0], out[0][0]) <= 2e-6
    assert rel_err(out[1][1], out[0][1]) <= 5e-6
    assert rel_err(out[1][3], out[0][3]) <= 2e-5


@pytest.mark.parametrize("variant", list(VARIANTS), ids=list(VARIANTS))
def test_fused_inverse_round_trip(gpu, restore, variant):
    """x ~ inv(fwd(x)) and lj_fwd ~ -lj_inv through the fused kernel in both directions (reference test/flow.jl:25-39)."""
    nf = gpu
    of32, of64 = _pair(64, [256, 256], 3)
    _set(nf, 1, VARIANTS[variant])
    gf = gpu_flow(nf, of32, np.float32)
    xs = z0(777, 64, np.float32, seed=9)
    y, lj = gf.with_logabsdet_jacobian(xs)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    np.testing.assert_allclose(xr, xs, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(lj, -lji, rtol=2e-5, atol=2e-5)
    x64, lji64 = of64.inverse(torch.from_numpy(y).double())
    assert rel_err(xr, x64.detach().numpy()) <= 5e-6
    lp = gf.logpdf(y)
    lp64 = of64.logpdf(torch.from_numpy(y).double()).detach().numpy()
    np.testing.assert_allclose(lp, lp64, rtol=2e-5, atol=2e-4)


def test_fused_variants_agree_at_scale(gpu, restore):
    """Both kernels, many tiles per CTA (persistent loop, tile hand-over, stash consumed by the backward pass)."""
    nf = gpu
    of32, _ = _pair(64, [256, 256], 2)
    gt = gpu_target(nf, oracle_target("funnel", 64))
    xs = z0(148 * 128 * 3 + 77, 64, np.float32, seed=3)
    res = {}
    for name, var in VARIANTS.items():
        _set(nf, 1, var)
        gf = gpu_flow(nf, of32, np.float32)
        res[name] = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    _set(nf, 0, 0)
    gf = gpu_flow(nf, of32, np.float32)
    v0, g0 = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    for name, (v, g) in res.items():
        assert abs(v - v0) <= 5e-6 * max(abs(v0), 1.0), name
        assert rel_err(g, g0) <= 2e-5, name


@pytest.mark.parametrize("variant", list(VARIANTS), ids=list(VARIANTS))
def test_fused_state_is_reproducible(gpu, restore, variant):
    """No atomics on the path of the new state (nor, in the two-team kernel, of the logdet): two runs over many tiles per CTA
    must agree bit for bit -- a missed hand-over between the epilogue warps, the tile warps and the tensor pipe shows up here."""
    nf = gpu
    of32, _ = _pair(64, [256, 256], 2)
    _set(nf, 1, VARIANTS[variant])
    gf = gpu_flow(nf, of32, np.float32)
    xs = z0(148 * 128 * 4 + 5, 64, np.float32, seed=11)
    y1, ld1 = gf.with_logabsdet_jacobian(xs)
    for _ in range(3):
        y2, ld2 = gf.with_logabsdet_jacobian(xs)
        assert np.array_equal(y1, y2)
        if variant == "two_team":
            assert np.array_equal(ld1, ld2)
        else:       # four shared-memory atomic adds per row: the order of the partial sums is free
            np.testing.assert_allclose(ld1, ld2, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("variant", list(VARIANTS), ids=list(VARIANTS))
def test_fused_no_stash_mode_gives_the_same_state(gpu, restore, variant):
    """Calls that no backward pass follows (forward, sampling, logpdf) skip every stash store inside the fused kernel: the state and
    logdet must be bit-identical to the stashing call (`nf_forward_stash`), and the stash of the latter must still drive
    `nf_backward` to the fused ELBO gradient."""
    nf = gpu
    of32, _ = _pair(64, [256, 256], 2)
    _set(nf, 1, VARIANTS[variant])
    gf = gpu_flow(nf, of32, np.float32)
    xs = z0(1500, 64, np.float32, seed=21)
    y_plain, ld_plain = gf.with_logabsdet_jacobian(xs)              # no stash
    y_stash, ld_stash = nf.forward_stash(gf, xs)                    # stash kept for nf_backward
    assert np.array_equal(y_plain, y_stash)
    if variant == "two_team":
        assert np.array_equal(ld_plain, ld_stash)
    else:
        np.testing.assert_allclose(ld_plain, ld_stash, rtol=1e-5, atol=1e-5)
    rng = np.random.Generator(np.random.PCG64(4))
    mu, sg = rng.standard_normal(64), rng.uniform(0.5, 1.5, 64)
    score = (-(y_stash - mu) / sg ** 2).astype(np.float32)
    g2 = nf.backward(gf, score / len(xs), np.full(len(xs), 1.0 / len(xs), dtype=np.float32))
    v, g = nf.api._elbo_impl(gf, nf.DiagNormal(mu, sg), xs, want_grad=True)
    assert np.linalg.norm(g2 - g) <= 1e-4 * np.linalg.norm(g)
