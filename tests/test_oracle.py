"""Pins the CPU oracle with the reference's own property tests plus autograd-vs-finite-
difference checks.  The reference carries no golden vectors, so this is all the pinning that exists
("parity unpinned" for fixed values -- see oracle/nf_oracle.py header and DESIGN.md)."""
import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import oracle_flow, oracle_target

DT = [torch.float32, torch.float64]


@pytest.mark.parametrize("dt", DT)
def test_analytic_elbo_is_zero(dt):
    """reference test/objectives.jl:1-26."""
    rng = np.random.Generator(np.random.PCG64(1))
    mu = rng.standard_normal(2); var = rng.random(2) + 1e-3
    target = O.DiagNormal(mu, np.sqrt(var))
    flow = O.shift_scale_flow(mu, np.sqrt(var), dt, base_sigma=[1.0, 1.0])
    xs = torch.from_numpy(rng.standard_normal((10, 2))).to(dt)
    for fn in (O.elbo, O.elbo_batch):
        el = float(fn(flow, target, xs))
        assert abs(el) <= 1e-5
    x = torch.from_numpy(rng.standard_normal((1, 2))).to(dt)
    lhs = float(flow.logpdf(x)[0]) + float(O.elbo(flow, target, xs))
    assert lhs == pytest.approx(float(target.logp(x)[0]), rel=1e-5, abs=1e-5)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("kind,rtol", [("realnvp", 1e-6), ("nsf", 1e-4), ("planar", 1e-4), ("radial", 1e-4)])
def test_inverse_consistency(kind, rtol, dt):
    """reference test/flow.jl:25-39,92-106,158-172,224-238 (d = 5: uneven masks 3/2)."""
    npdt = np.float32 if dt == torch.float32 else np.float64
    f = oracle_flow(kind, 5, npdt)
    if dt == torch.float32:
        rtol = max(rtol, 2e-4 if kind in ("planar",) else 1e-5)
    g = torch.Generator().manual_seed(0)
    for n in (1, 10):
        x = torch.randn(n, 5, generator=g, dtype=dt)
        y, lj = f.forward(x)
        xr, lji = f.inverse(y)
        assert torch.allclose(x, xr, rtol=rtol, atol=rtol)
        assert torch.allclose(lj, -lji, rtol=rtol, atol=rtol)


@pytest.mark.parametrize("kind", ["realnvp", "nsf", "planar", "radial"])
def test_elbo_finite_and_elbo_equals_elbo_batch(kind):
    """reference test/flow.jl:42-61; elbo (per column) and elbo_batch are the same number."""
    f = oracle_flow(kind, 5, np.float64)
    tgt = oracle_target("diag", 5)
    g = torch.Generator().manual_seed(1)
    for n in (64, 1):
        xs = torch.randn(n, 5, generator=g, dtype=torch.float64)
        a, b = float(O.elbo(f, tgt, xs)), float(O.elbo_batch(f, tgt, xs))
        assert np.isfinite(a) and np.isfinite(b)
        assert a == pytest.approx(b, rel=1e-12, abs=1e-12)


def test_shift_scale_training_converges():
    """reference test/interface.jl:1-53 with the oracle's Adam (Optimisers.Adam restatement)."""
    target = O.DiagNormal(10 * np.ones(2), 2 * np.ones(2))
    flow = O.shift_scale_flow(np.zeros(2), np.ones(2), torch.float64, base_sigma=[1.0, 1.0])
    theta = flow.theta().numpy()
    opt = O.Adam(1e-2)
    rng = np.random.Generator(np.random.PCG64(0))
    for it in range(5000):
        xs = torch.from_numpy(rng.standard_normal((10, 2)))
        v, g = O.elbo_value_and_grad(flow, target, theta, xs)
        if np.linalg.norm(g) < 1e-3:
            break
        theta = opt.update(theta, -g)          # loss = -elbo
    assert np.all(np.abs(theta[:2] - 10) < 0.2)
    assert np.all(np.abs(theta[2:] - 2) < 0.2)
    flow.set_theta(torch.from_numpy(theta))
    assert float(O.elbo(flow, target, torch.from_numpy(rng.standard_normal((1000, 2))))) > -1


def test_loglikelihood_prefers_trained_samples():
    """reference test/objectives.jl:28-35."""
    rng = np.random.Generator(np.random.PCG64(2))
    mu = rng.standard_normal(2); sd = np.sqrt(rng.random(2) + 1e-3)
    flow = O.shift_scale_flow(mu, sd, torch.float64, base_sigma=[1.0, 1.0])
    z = torch.from_numpy(rng.standard_normal((1000, 2)))
    y, _ = flow.forward(z)
    assert float(O.loglikelihood(flow, y)) > float(O.loglikelihood(flow, z))


def test_rqs_properties():
    """App. A.4: monotone, identity with logJ = 0 outside [-B, B], derivative 1 at the boundary knots."""
    rng = np.random.Generator(np.random.PCG64(3))
    K, c, B = 10, 3, 5.0
    raw = torch.from_numpy(rng.standard_normal((1, (3 * K - 1) * c)))
    pX, pY, dYdX = O.rqs_params_from_nn(raw, c, B)
    assert torch.allclose(pX[..., 0], torch.tensor(-B, dtype=torch.float64)) and torch.allclose(pX[..., -1], torch.tensor(B, dtype=torch.float64), atol=1e-12)
    assert torch.all(pX[..., 1:] > pX[..., :-1]) and torch.all(pY[..., 1:] > pY[..., :-1])
    assert torch.all(dYdX[..., 0] == 1) and torch.all(dYdX[..., -1] == 1)
    xs = torch.linspace(-7, 7, 2001, dtype=torch.float64)
    X = xs[:, None].expand(-1, c)
    P = lambda t: t.expand(xs.numel(), -1, -1)
    y, lj, k = O.rqs_forward(X, P(pX), P(pY), P(dYdX))
    assert torch.all(y[1:] >= y[:-1])                                  # monotone
    out = (xs < -B) | (xs > B)
    assert torch.equal(y[out], X[out]) and torch.all(lj[out] == 0)     # identity tails
    xi, lji, _ = O.rqs_inverse(y, P(pX), P(pY), P(dYdX))
    assert torch.allclose(xi, X, atol=1e-9) and torch.allclose(lj, -lji, atol=1e-8)
    eps = 1e-6                                                         # slope -> 1 at +-B from inside
    for b in (-B + eps, B - eps):
        xb = torch.full((1, c), b, dtype=torch.float64)
        _, ljb, _ = O.rqs_forward(xb, pX, pY, dYdX)
        assert torch.allclose(ljb, torch.zeros_like(ljb), atol=1e-4)


def test_bin_convention():
    """searchsortedfirst - 1: bins are (pX[k], pX[k+1]]; a value exactly on a knot belongs to the bin on its left."""
    knots = torch.tensor([[-1.0, 0.0, 1.0]])
    assert O.rqs_bin_index(knots, torch.tensor([0.0])).item() == 1
    assert O.rqs_bin_index(knots, torch.tensor([-1.0])).item() == 0
    assert O.rqs_bin_index(knots, torch.tensor([1.0])).item() == 2
    assert O.rqs_bin_index(knots, torch.tensor([1.5])).item() == 3


@pytest.mark.parametrize("kind,dim,tname,kw", [
    ("planar", 3, "banana", dict(nlayers=4)), ("radial", 3, "diag", dict(nlayers=4)),
    ("realnvp", 5, "funnel", dict(hdims=[8, 8], nlayers=1)), ("nsf", 4, "cross", dict(hdims=[8, 8], K=5, B=3.0, nlayers=1)),
    ("radial", 2, "warped", dict(nlayers=3))])
def test_autograd_matches_finite_differences(kind, dim, tname, kw):
    f = oracle_flow(kind, dim, np.float64, **kw)
    tgt = oracle_target(tname, dim)
    g = torch.Generator().manual_seed(5)
    xs = torch.randn(7, dim, generator=g, dtype=torch.float64)
    theta = f.theta().numpy()
    v, grad = O.elbo_value_and_grad(f, tgt, theta, xs)
    rng = np.random.Generator(np.random.PCG64(9))
    for i in rng.choice(theta.size, size=min(12, theta.size), replace=False):
        h = 1e-6 * max(1.0, abs(theta[i]))
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h; tm[i] -= h
        f.set_theta(torch.from_numpy(tp)); vp = float(O.elbo_batch(f, tgt, xs))
        f.set_theta(torch.from_numpy(tm)); vm = float(O.elbo_batch(f, tgt, xs))
        fd = (vp - vm) / (2 * h)
        assert fd == pytest.approx(grad[i], rel=2e-4, abs=1e-6), (i, fd, grad[i])
    f.set_theta(torch.from_numpy(theta))


def test_destructure_order_and_counts():
    """App. A.7: theta = [theta(Ls[1]); ...; theta(Ls[n])]; P of BASELINE config 3 is 1 319 424."""
    f = O.realnvp(64, [256, 256], 4, torch.float32)
    assert f.n_params() == 1319424
    assert O.nsf(16, [32, 32], 10, 5.0, 4, torch.float32).n_params() == 72000
    f = O.shift_scale_flow([1.0, 2.0], [3.0, 4.0])
    assert f.theta().tolist() == [1.0, 2.0, 3.0, 4.0]          # Shift first, then Scale (test/interface.jl:47-48)
    f = O.planarflow(2, 3)
    th = f.theta()
    assert torch.equal(th[:2], f.layers[0].w) and torch.equal(th[2:4], f.layers[0].u) and th[4] == f.layers[0].b[0]


def test_targets_match_closed_forms():
    y = torch.tensor([[0.3, -1.2]], dtype=torch.float64)
    # Banana(2, b=1, v=10): phi^-1(y) = (y1, y2 + b*y1^2 - v*b)
    u2 = -1.2 + 0.09 - 10
    ref = -0.5 * (np.log(10) + 2 * np.log(2 * np.pi)) - 0.5 * (0.09 / 10 + u2 * u2)
    assert float(O.Banana(2, 1.0, 10.0).logp(y)[0]) == pytest.approx(ref, rel=1e-12)
    # Funnel(2, 0, 9)
    ref = (-0.5 * np.log(2 * np.pi) - np.log(9) - 0.09 / 162) + (-0.5 * (np.log(2 * np.pi) + 0.3) - 0.5 * np.exp(-0.3) * 1.44)
    assert float(O.Funnel(2).logp(y)[0]) == pytest.approx(ref, rel=1e-12)


# ---- Hamiltonian flow (reference example/demo_hamiltonian_flow.jl) ---------------------------------------------
def _ham(tgt, nlayers=4, L=3, seed=0):
    f = O.hamiltonian_flow(tgt, nlayers, L, np.log(0.05))
    rng = np.random.default_rng(seed)
    th = f.theta().numpy()
    f.set_theta(torch.from_numpy(th + 0.05 * rng.standard_normal(th.size)))
    return f


def test_leapfrog_is_reversible_and_volume_preserving():
    """demo_hamiltonian_flow.jl:63-91: inverse = leapfrog with -eps; logabsdetjac = 0 (checked against autograd's Jacobian)."""
    tgt = O.Funnel(2, -2.0, 3.0)
    lf = O.LeapFrog(torch.tensor([-2.0, -2.5], dtype=torch.float64), 3, O.target_score(tgt))
    z = torch.from_numpy(np.random.default_rng(1).standard_normal((5, 4)))
    y, ld = lf.forward(z)
    zr, _ = lf.inverse(y)
    assert float((zr - z).abs().max()) < 1e-12 and float(ld.abs().max()) == 0.0
    J = torch.autograd.functional.jacobian(lambda v: lf.forward(v[None, :])[0][0], z[0])
    assert abs(float(torch.linalg.det(J)) - 1.0) < 1e-10


def test_target_scores_match_autograd():
    rng = np.random.default_rng(2)
    for tgt in (O.Funnel(3, -1.0, 2.0), O.Banana(3, 0.4, 5.0), O.DiagNormal(rng.standard_normal(3), rng.uniform(0.5, 2, 3))):
        x = torch.from_numpy(rng.standard_normal((7, 3))).requires_grad_(True)
        (g,) = torch.autograd.grad(tgt.logp(x).sum(), x)
        assert float((g - O.target_score(tgt)(x.detach())).abs().max()) < 1e-12


def test_hamiltonian_flow_theta_layout_and_gradient():
    """theta = per layer [b; a; log_eps], then q0's shift, scale; autograd gradient vs central differences."""
    tgt = O.Funnel(2, -2.0, 3.0)
    f = _ham(tgt)
    assert f.n_params() == 4 * 6 + 8
    jt = O.JointTarget(tgt)
    xs = torch.from_numpy(np.random.default_rng(3).standard_normal((32, 4)))
    th = f.theta().numpy().copy()
    v, g = O.elbo_value_and_grad(f, jt, th, xs)
    for i in (0, 3, 5, 11, 24, 30):
        e = np.zeros_like(th); e[i] = 1e-6
        vp, _ = O.elbo_value_and_grad(f, jt, th + e, xs)
        vm, _ = O.elbo_value_and_grad(f, jt, th - e, xs)
        assert abs((vp - vm) / 2e-6 - g[i]) < 1e-6 * max(1.0, abs(g[i]))
    # joint target = target + standard normal momentum
    z = xs[:3]
    ref = tgt.logp(z[:, :2]) + torch.distributions.Normal(0, 1).log_prob(z[:, 2:]).sum(dim=1)
    assert float((jt.logp(z) - ref).abs().max()) < 1e-12


def test_logreg_target_score_and_normalisation():
    """LogReg: closed-form score == autograd; with no data the target is exactly N(0, sigma0^2 I)."""
    tgt = O.synthetic_logreg(5, 30, sigma0=1.5)
    x = torch.from_numpy(np.random.default_rng(4).standard_normal((6, 5))).requires_grad_(True)
    (g,) = torch.autograd.grad(tgt.logp(x).sum(), x)
    assert float((g - tgt.score(x.detach())).abs().max()) < 1e-12
    empty = O.LogReg(np.zeros((0, 5)), np.zeros(0), sigma0=1.5)
    ref = torch.distributions.Normal(torch.tensor(0.0, dtype=torch.float64), torch.tensor(1.5, dtype=torch.float64)).log_prob(x.detach()).sum(dim=1)
    assert float((empty.logp(x.detach()) - ref).abs().max()) < 1e-12


@pytest.mark.parametrize("kind,kw", [("planar", dict(nlayers=3)), ("radial", dict(nlayers=3)),
                                     ("realnvp", dict(hdims=[8, 8], nlayers=1)),
                                     ("nsf", dict(hdims=[8, 8], K=6, B=3.0, nlayers=1))])
def test_logdet_equals_log_abs_det_of_autograd_jacobian(kind, kw):
    """Independent pin of every layer family's hand-written log|det J| (Bijectors / MonotonicSplines formulas restated in the
    oracle): it must equal log|det| of the Jacobian torch.autograd builds from the forward map alone."""
    from helpers import oracle_flow
    of = oracle_flow(kind, 4, np.float64, **kw)
    xs = torch.from_numpy(np.random.default_rng(12).standard_normal((5, 4)) * 1.5)
    _, ld = of.forward(xs)
    for i in range(xs.shape[0]):
        J = torch.autograd.functional.jacobian(lambda v: of.forward(v[None, :])[0][0], xs[i])
        assert abs(float(torch.linalg.slogdet(J)[1]) - float(ld[i])) < 1e-9, (kind, i)


def test_dense_base_matches_scipy():
    """Full-covariance base (App. A.8): logpdf against scipy's multivariate_normal, sampling = mu + L z."""
    from scipy.stats import multivariate_normal
    rng = np.random.Generator(np.random.PCG64(5))
    d = 6
    A = rng.standard_normal((d, d))
    Sigma = A @ A.T + 0.5 * np.eye(d)
    mu = rng.standard_normal(d)
    f = O.shift_scale_flow(np.zeros(d), np.ones(d))
    f.base_mu = torch.from_numpy(mu)
    f.base_chol = torch.from_numpy(np.linalg.cholesky(Sigma))
    x = rng.standard_normal((50, d))
    ref = multivariate_normal(mu, Sigma).logpdf(x)
    assert np.allclose(f.base_logpdf(torch.from_numpy(x)).numpy(), ref, rtol=1e-12, atol=1e-12)
    z = rng.standard_normal((200000, d))
    xs = f.base_sample(torch.from_numpy(z)).numpy()
    assert np.abs(np.cov(xs.T) - Sigma).max() < 0.1 and np.abs(xs.mean(0) - mu).max() < 0.03
