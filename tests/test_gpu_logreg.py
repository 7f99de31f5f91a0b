"""Logistic-regression posterior target (BASELINE config 5's target, synthetic) on every flow family,
including as the LeapFrog score (second-order terms = analytic Hessian-vector products), against the oracle."""
import math

import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, rel_err, z0

pytestmark = pytest.mark.gpu

TDT = {np.float32: torch.float32, np.float64: torch.float64}
TOL = {np.float32: (1e-5, 1e-4), np.float64: (1e-9, 1e-7)}


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind,dim,kw", [("planar", 4, dict(nlayers=6)), ("radial", 8, dict(nlayers=4)),
                                         ("realnvp", 10, dict(hdims=[32, 32], nlayers=2)),
                                         ("nsf", 6, dict(hdims=[16, 16], K=8, B=4.0, nlayers=1)),
                                         ("realnvp", 100, dict(hdims=[64, 64], nlayers=1))],     # BASELINE's 100-D posterior
                         ids=["planar-d4", "radial-d8", "realnvp-d10", "nsf-d6", "realnvp-d100"])
def test_logreg_target_elbo_value_and_grad(gpu, kind, dim, kw, dtype):
    nf = gpu
    of = oracle_flow(kind, dim, dtype, **kw)
    ot = O.synthetic_logreg(dim, 60)
    xs = z0(200, dim, dtype)
    v_ref, g_ref = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs))
    tv, tg = TOL[dtype]
    if dtype == np.float32:
        of64 = oracle_flow(kind, dim, np.float64, **kw)
        of64.set_theta(of.theta().double())
        v64, g64 = O.elbo_value_and_grad(of64, ot, of64.theta(), torch.from_numpy(xs).double())
        tv, tg = max(tv, 2 * abs(v_ref - v64) / max(abs(v64), 1.0)), max(tg, 2 * rel_err(g_ref, g64))
    v, g = nf.api._elbo_impl(gpu_flow(nf, of, dtype), gpu_target(nf, ot), xs, want_grad=True)
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("h", [2, 8])
def test_hamiltonian_flow_on_logreg_posterior(gpu, h, dtype):
    """LeapFrog driven by the logistic-regression score; joint target logp(beta) + logN(rho)."""
    nf = gpu
    tgt = O.synthetic_logreg(h, 40)
    of = O.hamiltonian_flow(tgt, 3, 2, math.log(0.05), dtype=TDT[dtype])
    rng = np.random.default_rng(0)
    th = of.theta().double().numpy()
    of.set_theta(torch.from_numpy(th + 0.05 * rng.standard_normal(th.size)).to(TDT[dtype]))
    jt = O.JointTarget(tgt)
    xs = z0(128, 2 * h, dtype)
    v_ref, g_ref = O.elbo_value_and_grad(of, jt, of.theta(), torch.from_numpy(xs))
    tv, tg = TOL[dtype]
    if dtype == np.float32:
        tv, tg = 2e-5, 2e-4
    v, g = nf.api._elbo_impl(gpu_flow(nf, of, dtype), gpu_target(nf, jt), xs, want_grad=True)
    assert abs(v - v_ref) <= tv * max(abs(v_ref), 1.0), (v, v_ref)
    assert rel_err(g, g_ref) <= tg, rel_err(g, g_ref)
    # reversible + volume preserving with this score too
    gf = gpu_flow(nf, of, dtype)
    y, lj = gf.with_logabsdet_jacobian(xs)
    xr, lji = gf.inverse_with_logabsdet_jacobian(y)
    rt = 1e-4 if dtype == np.float32 else 1e-10
    np.testing.assert_allclose(xr, xs, rtol=rt, atol=rt)
    np.testing.assert_allclose(lj, -lji, rtol=rt, atol=rt)


def test_logreg_training_improves_elbo(gpu):
    """RealNVP on a 16-D logistic-regression posterior: a few hundred Adam steps must raise the ELBO."""
    nf = gpu
    nf.seed(3)
    ot = O.synthetic_logreg(16, 200)
    tgt = nf.LogReg(ot.X.numpy(), ot.y.numpy(), ot.sigma0)
    flow = nf.realnvp(nf.MvNormal(np.zeros(16)), [32, 32], 2, np.float32)
    _, stats, _ = nf.train_flow(np.random.default_rng(1), nf.elbo_batch, flow, tgt, 2048, max_iters=200, optimiser=nf.Adam(2e-3),
                                ADbackend=nf.AutoNFCUDA(), show_progress=False)
    l0 = np.mean([s["loss"] for s in stats[:10]]); l1 = np.mean([s["loss"] for s in stats[-10:]])
    assert np.isfinite(l1) and l1 < l0 - 1.0, (l0, l1)
