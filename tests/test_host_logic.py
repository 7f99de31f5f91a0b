"""Host-side mirror of the reference interface: structure, destructure layout, optimiser (no GPU needed)."""
import numpy as np
import pytest
import torch

import nf_oracle as O


def test_constructors_match_reference_structure(nf):
    q0 = nf.MvNormal(np.zeros(64))
    f = nf.realnvp(q0, [256, 256], 4, nf.Float32)                       # src/flows/realnvp.jl:170-180
    assert f.num_params == 1319424 and len(f.layers) == 8 and f.theta.dtype == np.float32
    assert f.layers[0].mask_idx == list(range(0, 64, 2)) and f.layers[1].mask_idx == list(range(1, 64, 2))
    g = nf.nsf(nf.MvNormal(np.zeros(16)), [32, 32], 10, 5.0, 4, nf.Float32)   # src/flows/neuralspline.jl:218-230
    assert g.num_params == 72000 and g.layers[0].K == 10 and g.layers[0].B == 5.0
    assert nf.planarflow(nf.MvNormal(np.zeros(2)), 20).num_params == 100
    assert nf.radialflow(nf.MvNormal(np.zeros(2)), 20).num_params == 80
    d = nf.realnvp(nf.MvNormal(np.zeros(5)), [32, 32], 2)                    # odd d: masks of 3 and 2 (test/flow.jl:4)
    assert [len(l.mask_idx) for l in d.layers] == [3, 2, 3, 2]
    assert nf.realnvp(nf.MvNormal(np.zeros(2))).num_params == O.realnvp(2, [32, 32], 10).n_params()   # defaults :190-192


def test_destructure_and_re(nf):
    flow = nf.transformed(nf.MvNormal(np.zeros(2), np.ones(2)), nf.Shift([1.0, 2.0]) @ nf.Scale([3.0, 4.0]))
    theta, re = nf.destructure(flow)
    assert theta.tolist() == [1.0, 2.0, 3.0, 4.0]                          # Shift first, then Scale (test/interface.jl:47-48)
    g = re(theta * 2)
    assert g.theta.tolist() == [2.0, 4.0, 6.0, 8.0] and flow.theta.tolist() == [1.0, 2.0, 3.0, 4.0]
    with pytest.raises(ValueError):
        re(np.zeros(3))


def test_layer_theta_layout_matches_oracle(nf):
    """Same per-layer parameter counts and ordering as the oracle's destructure restatement."""
    nf.seed(7)
    f = nf.create_flow([nf.RealNVP_layer(5, [8, 8]), nf.NSF_layer(5, [8, 8], 4, 3.0)], nf.MvNormal(np.zeros(5)))
    of = O.Flow(5, O.realnvp(5, [8, 8], 1).layers + O.nsf(5, [8, 8], 4, 3.0, 1).layers)
    assert f.num_params == of.n_params()
    assert [l.theta.size for l in f.layers] == [l.n_params() for l in of.layers]
    # Dense biases start at zero, weights inside the glorot bound (Flux, App. A.6)
    aff = f.layers[0]
    w1 = aff.theta[:2 * 8]
    assert np.all(np.abs(w1) <= np.sqrt(6 / (2 + 8)) + 1e-7) and np.all(aff.theta[16:24] == 0)


def test_adam_matches_optimisers_rule(nf):
    rng = np.random.Generator(np.random.PCG64(0))
    theta = rng.standard_normal(10)
    a, b = nf.Adam(1e-2), O.Adam(1e-2)
    st = a.setup(theta)
    ta, tb = theta.copy(), theta.copy()
    for _ in range(20):
        g = rng.standard_normal(10)
        st, ta = a.update(st, ta, g)
        tb = b.update(tb, g)
    assert np.allclose(ta, tb, rtol=1e-12, atol=1e-12)


def test_train_flow_requires_adbackend(nf):
    flow = nf.planarflow(nf.MvNormal(np.zeros(2)), 2)
    with pytest.raises(TypeError):
        nf.train_flow(nf.elbo, flow, nf.Banana(2, 1.0, 10.0), 10)


def test_argument_validation(nf):
    with pytest.raises(ValueError):
        nf.Banana(1, 1.0, 1.0)
    with pytest.raises(ValueError):
        nf.Funnel(2, 0.0, -1.0)
    with pytest.raises(ValueError):
        nf.MvNormal([0.0, 0.0], [1.0, -1.0])
    with pytest.raises(ValueError):
        nf.Flow([nf.PlanarLayer(3)], nf.MvNormal(np.zeros(2)))
    with pytest.raises(TypeError):
        nf.realnvp(nf.MvNormal(np.zeros(2)), [8], 1, np.int32)


def test_shard_ranges(nf):
    for n, w in [(10, 3), (1 << 20, 8), (5, 8), (7, 1)]:
        spans = [nf.dp.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
