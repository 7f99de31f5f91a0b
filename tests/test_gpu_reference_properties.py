"""The reference's own property tests, run against the CUDA path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
def test_objectives_analytic_elbo(gpu, T):
    """reference test/objectives.jl:1-37: the flow is the exact affine map of the target -> ELBO == 0."""
    nf = gpu
    rng = np.random.Generator(np.random.PCG64(1))
    mu = rng.standard_normal(2).astype(T)
    var = (rng.random(2) + 1e-3).astype(T)
    target = nf.DiagNormal(mu, np.sqrt(var))
    q0 = nf.MvNormal(np.zeros(2), np.ones(2))
    flow = nf.transformed(q0, nf.Shift(mu) @ nf.Scale(np.sqrt(var)), T)
    el = nf.elbo(rng, flow, target, 10)
    assert abs(el) <= 1e-5
    elb = nf.elbo_batch(rng, flow, target, 10)
    assert abs(elb) <= 1e-5


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["realnvp", "nsf", "planar", "radial"])
def test_flow_elbo_finite(gpu, kind, T):
    """reference test/flow.jl:42-61: elbo / elbo_batch finite at n = 64 and n = 1 (d = 5)."""
    nf = gpu
    dim = 5
    q0 = nf.MvNormal(np.zeros(dim))
    nf.seed(123)
    flow = {"realnvp": lambda: nf.realnvp(q0, [32, 32], 2, T), "nsf": lambda: nf.nsf(q0, [32, 32], 10, 5.0, 2, T),
            "planar": lambda: nf.planarflow(q0, 10, T), "radial": lambda: nf.radialflow(q0, 10, T)}[kind]()
    rng = np.random.Generator(np.random.PCG64(2))
    target = nf.DiagNormal(rng.standard_normal(dim), np.sqrt(rng.random(dim) + 1e-3))
    assert np.isfinite(nf.elbo(rng, flow, target, 64))
    assert np.isfinite(nf.elbo_batch(rng, flow, target, 64))
    assert np.isfinite(nf.elbo(rng, flow, target, 1))
    if kind in ("realnvp", "nsf"):
        ys = flow.rand(100)
        assert ys.shape == (100, dim) and ys.dtype == T
        ls = flow.logpdf(ys)
        assert ls.shape == (100,) and ls.dtype == T and np.all(np.isfinite(ls))


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
def test_interface_train_flow_converges(gpu, T):
    """reference test/interface.jl:1-53: train_flow(elbo, Shift∘Scale flow) -> theta ≈ (10,10,2,2)."""
    nf = gpu
    mu = 10 * np.ones(2)
    target = nf.DiagNormal(mu, 2 * np.ones(2))
    q0 = nf.MvNormal(np.zeros(2), np.ones(2))
    flow = nf.transformed(q0, nf.Shift(np.zeros(2)) @ nf.Scale(np.ones(2)), T)
    seen = {"cb": 0}

    def cb(it, opt_stats, re, theta):
        seen["cb"] += 1
        return {"sample_per_iter": 10}

    def checkconv(it, stat, re, theta, st):
        return stat["gradient_norm"] < 1e-3

    rng = np.random.Generator(np.random.PCG64(0))
    flow_trained, stats, _ = nf.train_flow(rng, nf.elbo, flow, target, 10, max_iters=5000, optimiser=nf.Adam(0.01),
                                           ADbackend=nf.AutoNFCUDA(), show_progress=False, callback=cb, hasconverged=checkconv)
    theta, re = nf.destructure(flow_trained)
    el_untrained = nf.elbo(rng, flow, target, 1000)
    el_trained = nf.elbo(rng, flow_trained, target, 1000)
    assert np.all(np.abs(theta[:2] - mu) < 0.2)
    assert np.all(np.abs(theta[2:] - 2) < 0.2)
    assert el_trained > el_untrained
    assert el_trained > -1
    assert seen["cb"] == len(stats) and "sample_per_iter" in stats[0]
    with pytest.raises(TypeError):
        nf.train_flow(nf.elbo, flow, target, 10)     # ADbackend is required, as in the reference


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
def test_two_phase_api_matches_fused_elbo(gpu, T):
    """nf_forward_stash + caller-side logp/score + nf_backward == the fused ELBO gradient (user-supplied log-density path)."""
    nf = gpu
    dim = 6
    nf.seed(5)
    flow = nf.realnvp(nf.MvNormal(np.zeros(dim)), [16, 16], 2, T)
    rng = np.random.Generator(np.random.PCG64(4))
    mu, sg = rng.standard_normal(dim), rng.uniform(0.5, 1.5, dim)
    target = nf.DiagNormal(mu, sg)
    xs = rng.standard_normal((200, dim)).astype(T)
    v, g = nf.api._elbo_impl(flow, target, xs, want_grad=True)
    ys, ld = nf.forward_stash(flow, xs)
    score = (-(ys - mu) / sg ** 2).astype(T)                 # d logp / dy evaluated by the caller
    g2 = nf.backward(flow, score / len(xs), np.full(len(xs), 1.0 / len(xs), dtype=T))
    tol = 1e-4 if T == np.float32 else 1e-9
    assert np.linalg.norm(g2 - g) <= tol * np.linalg.norm(g)


def test_device_sampling_statistics(gpu):
    """_device_specific_rand replacement: Philox base draws ~ N(mu, sigma^2), reproducible per seed."""
    nf = gpu
    q0 = nf.MvNormal([1.0, -2.0, 0.5], [0.5, 2.0, 1.0])
    flow = nf.planarflow(q0, 2, np.float32)
    z = flow.rand_base(200000, seed=7)
    assert z.shape == (200000, 3) and z.dtype == np.float32
    assert np.allclose(z.mean(0), q0.mu, atol=0.02) and np.allclose(z.std(0), q0.sigma, rtol=0.02)
    assert np.array_equal(z, flow.rand_base(200000, seed=7)) and not np.array_equal(z, flow.rand_base(200000, seed=8))
    kurt = (((z - z.mean(0)) / z.std(0)) ** 4).mean(0)
    assert np.allclose(kurt, 3.0, atol=0.1)
    ys = flow.rand(1000, seed=3)
    assert ys.shape == (1000, 3) and np.all(np.isfinite(ys))


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["planar", "realnvp", "nsf"])
def test_on_device_adam_matches_host_loop(gpu, T, kind):
    """nf_train_elbo_adam == the reference loop body (value_and_gradient + Optimisers.Adam update) run from the host
    with the same per-iteration Philox seeds.  planar: the persistent single-launch loop; realnvp / nsf: one iteration captured
    in a CUDA graph and replayed (iteration index and seed offset read from a device counter)."""
    import ctypes as C
    nf = gpu
    K = nf._capi
    nf.seed(11)
    if kind == "planar":
        flow = nf.planarflow(nf.MvNormal(np.zeros(2)), 5, T)
    elif kind == "realnvp":
        flow = nf.realnvp(nf.MvNormal(np.zeros(2)), [16, 16], 2, T)
    else:
        flow = nf.nsf(nf.MvNormal(np.zeros(2)), [16, 16], 6, 4.0, 1, T)
    target = nf.Banana(2, 1.0, 10.0)
    n, iters, seed0 = 256, 25, 1234
    # host loop
    theta = flow.theta.copy()
    opt = nf.Adam(1e-2)
    st = opt.setup(theta)
    _, re = nf.destructure(flow)
    losses = []
    for i in range(iters):
        ls, g = nf.api._elbo_impl(re(theta), target, n, want_grad=True, scale=-1.0, seed=seed0 + i)
        losses.append(ls)
        st, theta = opt.update(st, theta, g)
    # device loop
    theta_d = flow.theta.copy()
    m = np.zeros_like(theta_d); v = np.zeros_like(theta_d)
    stats = np.empty((iters, 2))
    K.check(K.lib().nf_train_elbo_adam(flow.handle(), target.handle(), K.ptr(theta_d), n, seed0, iters, 0, 1e-2, 0.9, 0.999, 1e-8,
                                       K.ptr(m), K.ptr(v), stats.ctypes.data_as(C.POINTER(C.c_double))))
    tol = (2e-4 if kind == "planar" else 2e-3) if T == np.float32 else 1e-9   # 25 chained Float32 Adam steps through tensor-core MLPs
    assert np.allclose(theta_d, theta, rtol=tol, atol=tol)
    assert np.allclose(stats[:, 0], np.array(losses, dtype=np.float64), rtol=tol, atol=tol)
    assert np.allclose(m, st["m"], rtol=10 * tol, atol=tol)


def test_train_flow_on_device_converges(gpu):
    """reference test/interface.jl convergence criterion with the optimiser step kept on the GPU."""
    nf = gpu
    mu = 10 * np.ones(2)
    target = nf.DiagNormal(mu, 2 * np.ones(2))
    flow = nf.transformed(nf.MvNormal(np.zeros(2), np.ones(2)), nf.Shift(np.zeros(2)) @ nf.Scale(np.ones(2)), np.float32)
    rng = np.random.Generator(np.random.PCG64(0))
    flow_trained, stats, st = nf.train_flow(rng, nf.elbo, flow, target, 10, max_iters=3000, optimiser=nf.Adam(0.01),
                                            ADbackend=nf.AutoNFCUDA(on_device=True, chunk=500), show_progress=False)
    theta, _ = nf.destructure(flow_trained)
    assert len(stats) == 3000 and stats[0]["iteration"] == 1 and "gradient_norm" in stats[-1]
    assert np.all(np.abs(theta[:2] - mu) < 0.2) and np.all(np.abs(theta[2:] - 2) < 0.2)
    assert stats[-1]["loss"] < stats[0]["loss"]


@pytest.mark.parametrize("tname,dim", [("banana", 2), ("funnel", 64), ("warped", 2), ("cross", 16), ("diag", 5)])
def test_device_targets_match_reference_formulas(gpu, tname, dim):
    import torch
    from helpers import gpu_target, oracle_target, z0
    """nf_target_logp: the device log-densities and scores equal the restated reference formulas (example/targets/*.jl) --
    the check the Julia shim runs against the `logp` closure it is handed before it trusts a named target."""
    nf = gpu
    ot = oracle_target(tname, dim)
    gt = gpu_target(nf, ot)
    xs = 0.8 * z0(64, dim, np.float64, seed=12)
    x = torch.from_numpy(xs).requires_grad_(True)
    lp_ref = ot.logp(x)
    sc_ref, = torch.autograd.grad(lp_ref.sum(), x)
    lp, sc = gt.logp(xs, np.float64, with_score=True)
    np.testing.assert_allclose(lp, lp_ref.detach().numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(sc, sc_ref.numpy(), rtol=1e-8, atol=1e-9)
    lp32 = gt.logp(xs.astype(np.float32), np.float32)
    np.testing.assert_allclose(lp32, lp_ref.detach().numpy(), rtol=2e-5, atol=2e-5)


def test_rand_advances_the_rng(gpu):
    """rand(flow, n) draws a fresh batch per call (the reference advances its RNG); an explicit seed reproduces."""
    nf = gpu
    nf.seed(5)
    flow = nf.planarflow(nf.MvNormal(np.zeros(2)), 3, np.float32)
    a, b = flow.rand(64), flow.rand(64)
    assert not np.array_equal(a, b)
    assert np.array_equal(flow.rand(64, seed=9), flow.rand(64, seed=9))
    assert not np.array_equal(flow.rand_base(64), flow.rand_base(64))
