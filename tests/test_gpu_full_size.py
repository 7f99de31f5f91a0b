"""BASELINE-size (N = 2^20) checks through size-independent properties, plus ragged / tiny batch sizes against the oracle.

The oracle cannot run 2^20 samples of the headline flows in seconds, so at full size the CUDA path is checked against itself
through properties the objective must have:
  * additivity over samples: sums over the two halves of the batch equal the sum over the whole batch (value and gradient) --
    catches anything that depends on tile / chunk / CTA position;
  * mean of the per-sample terms (`_batched_elbos`, reference src/objectives/elbo.jl:65-70) equals the ELBO value;
  * chunked (small workspace) == unchunked;
  * forward -> inverse round trip and logdet antisymmetry (reference test/flow.jl:25-39);
  * a random subsample of the full-size per-sample terms matches the oracle evaluated on those rows.
"""
import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0

pytestmark = pytest.mark.gpu

N_FULL = 1 << 20

FULL = [("realnvp", 64, "funnel", dict(hdims=[256, 256], nlayers=4)),            # BASELINE config 3
        ("nsf", 16, "cross", dict(hdims=[32, 32], K=10, B=5.0, nlayers=4))]      # BASELINE config 4


def _setup(nf, kind, dim, tname, kw):
    of = oracle_flow(kind, dim, np.float32, **kw)
    ot = oracle_target(tname, dim)
    return of, ot, gpu_flow(nf, of, np.float32), gpu_target(nf, ot)


@pytest.mark.parametrize("kind,dim,tname,kw", FULL, ids=[c[0] for c in FULL])
def test_full_size_additivity_terms_and_chunking(gpu, kind, dim, tname, kw):
    nf = gpu
    of, ot, gf, gt = _setup(nf, kind, dim, tname, kw)
    xs = z0(N_FULL, dim, np.float32)
    v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    assert np.isfinite(v) and np.all(np.isfinite(g))
    # Tolerances are the north-star ones (1e-5 value, 1e-4 gradient): the per-tensor fp16 plane scales follow the batch maximum,
    # so a sample's arithmetic is not bit-identical between the whole batch and a part of it, and Funnel(64) at random
    # initialisation is dominated by a few huge terms (|ELBO| ~ 1e17) that amplify every rounding difference.
    TV, TG = 1e-5, 1e-4
    h = N_FULL // 2
    v1, g1 = nf.api._elbo_impl(gf, gt, xs[:h], want_grad=True)
    v2, g2 = nf.api._elbo_impl(gf, gt, xs[h:], want_grad=True)
    assert abs(0.5 * (float(v1) + float(v2)) - float(v)) <= TV * max(abs(v), 1.0)
    assert rel_err(0.5 * (g1.astype(np.float64) + g2), g) <= TG
    # an odd split (tiles of 128 rows straddle the cut; the last tile is ragged)
    k = 333_333
    va, ga = nf.api._elbo_impl(gf, gt, xs[:k], want_grad=True)
    vb, gb = nf.api._elbo_impl(gf, gt, xs[k:], want_grad=True)
    assert abs((k * float(va) + (N_FULL - k) * float(vb)) / N_FULL - float(v)) <= TV * max(abs(v), 1.0)
    assert rel_err((k * ga.astype(np.float64) + (N_FULL - k) * gb) / N_FULL, g) <= TG
    # per-sample terms
    terms = nf.batched_elbos(gf, gt, xs)
    assert terms.shape == (N_FULL,)
    assert abs(float(np.mean(terms, dtype=np.float64)) - float(v)) <= TV * max(abs(v), 1.0)
    # subsample against the oracle (Float64 oracle at the same Float32-rounded theta; Float32 noise floor of the CPU path allowed)
    rows = np.random.default_rng(5).choice(N_FULL, 512, replace=False)
    of64 = oracle_flow(kind, dim, np.float64, **kw)
    of64.set_theta(of.theta().double())
    sub = torch.from_numpy(xs[rows])
    t64 = O.batched_elbos(of64, ot, sub.double()).detach().numpy()
    t32 = O.batched_elbos(of, ot, sub).detach().numpy().astype(np.float64)
    floor = abs(t32.mean() - t64.mean()) / max(abs(t64.mean()), 1.0)
    assert abs(terms[rows].astype(np.float64).mean() - t64.mean()) <= max(1e-5, 2 * floor) * max(abs(t64.mean()), 1.0)
    # chunked evaluation (workspace limit far below what the whole batch needs)
    gf2 = gpu_flow(nf, of, np.float32).set_workspace_limit(6 << 30)
    vc, gc = nf.api._elbo_impl(gf2, gt, xs, want_grad=True)
    assert abs(float(vc) - float(v)) <= TV * max(abs(v), 1.0)
    assert rel_err(gc, g) <= TG


@pytest.mark.parametrize("kind,dim,tname,kw", FULL, ids=[c[0] for c in FULL])
def test_full_size_round_trip(gpu, kind, dim, tname, kw):
    nf = gpu
    of, ot, gf, gt = _setup(nf, kind, dim, tname, kw)
    x = z0(N_FULL, dim, np.float32, seed=9)
    y, lj = gf.with_logabsdet_jacobian(x)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    assert np.all(np.isfinite(y)) and np.all(np.isfinite(lj))
    # the reference tests Float32 round trips at rtol 1e-4 (test/flow.jl:25-39); allow the same on 2^20 x d numbers
    err = np.abs(xr - x) / (1.0 + np.abs(x))
    assert float(err.max()) <= 1e-3 and float(np.mean(err)) <= 1e-5
    assert float(np.max(np.abs(lj + lji) / (1.0 + np.abs(lj)))) <= 1e-3


RAGGED = [1, 2, 127, 129, 1025]


@pytest.mark.parametrize("N", RAGGED)
@pytest.mark.parametrize("kind,dim,tname,kw", [("realnvp", 5, "diag", dict(hdims=[32, 32], nlayers=2)),
                                               ("nsf", 5, "diag", dict(hdims=[32, 32], K=10, B=5.0, nlayers=2)),
                                               ("planar", 2, "banana", dict(nlayers=5)), ("radial", 3, "diag", dict(nlayers=5))],
                         ids=["realnvp", "nsf", "planar", "radial"])
def test_ragged_and_tiny_batches(gpu, kind, dim, tname, kw, N):
    """n = 1 and batch sizes that straddle the 128-row GEMM tiles / 128-pair spline tiles (reference test/flow.jl:42-61 uses n = 1 and 64)."""
    nf = gpu
    of = oracle_flow(kind, dim, np.float64, **kw)
    ot = oracle_target(tname, dim)
    xs = z0(N, dim, np.float64, seed=N)
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    for dtype, tv, tg in ((np.float64, 1e-9, 1e-7), (np.float32, 2e-5, 2e-4)):
        gf = gpu_flow(nf, of, dtype)
        v, g = nf.api._elbo_impl(gf, gpu_target(nf, ot), xs.astype(dtype), want_grad=True)
        assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (dtype, v, v_ref)
        assert rel_err(g, g_ref) <= tg, (dtype, rel_err(g, g_ref))
