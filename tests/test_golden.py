"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the float64 oracle).

CPU: the oracle still reproduces them (guards the oracle against drift).
GPU: the CUDA path through the C ABI matches them within the north_star tolerances."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    d = np.load(path, allow_pickle=False)
    meta = json.loads(str(d["meta"]))
    of = oracle_flow(meta["kind"], meta["dim"], np.float64, **meta["kw"])
    if "theta" in d.files:
        theta32 = d["theta"].astype(np.float32)
    else:
        theta32 = of.theta().numpy().astype(np.float32)
        assert hashlib.sha256(theta32.tobytes()).hexdigest() == str(d["theta_sha256"])
    of.set_theta(torch.from_numpy(theta32).double())
    return d, meta, of, theta32


def _check_grad(d, g, tol):
    if "grad" in d.files:
        assert rel_err(g, d["grad"]) <= tol
    else:
        assert rel_err(g[:4096], d["grad_head"]) <= tol
        assert abs(np.linalg.norm(g.astype(np.float64)) - d["grad_norm"]) <= tol * d["grad_norm"]
        proj = float(g.astype(np.float64) @ np.cos(np.arange(g.size) * 0.001))
        assert abs(proj - d["grad_proj"]) <= tol * d["grad_norm"]


def test_fixtures_exist():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    d, meta, of, _ = _load(path)
    ot = oracle_target(meta["target"], meta["dim"])
    xs = torch.from_numpy(d["z0"]).double()
    v, g = O.elbo_value_and_grad(of, ot, of.theta(), xs)
    assert v == pytest.approx(float(d["elbo"]), rel=1e-12)
    _check_grad(d, g, 1e-10)
    y, ld = of.forward(xs)
    assert np.allclose(y.detach().numpy(), d["y"], rtol=1e-12, atol=1e-12)
    if "bins" in d.files:
        bins = np.stack([l.last_bins.numpy() for l in reversed(of.layers)])
        assert np.array_equal(bins, d["bins"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_matches_golden(gpu, path):
    nf = gpu
    d, meta, of, theta32 = _load(path)
    gf = gpu_flow(nf, of, np.float32)
    gf.theta = theta32
    gt = gpu_target(nf, oracle_target(meta["target"], meta["dim"]))
    xs = d["z0"]
    v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    assert abs(v - float(d["elbo"])) <= 1e-5 * max(abs(float(d["elbo"])), 1.0)
    _check_grad(d, g, 1e-4)
    y, ld = gf.with_logabsdet_jacobian(xs)
    assert rel_err(y, d["y"]) <= 1e-5 and rel_err(ld, d["logdet"]) <= 1e-4
    terms = nf.batched_elbos(gf, gt, xs)
    assert rel_err(terms, d["terms"]) <= 1e-5
    if "bins" in d.files:
        got = nf.spline_bins(gf, xs)
        # bins are integers: exact, except where the searched value sits within float32 rounding of the knot that separates
        # the two answers.  The fixture's bins come from the float64 oracle at this theta; every mismatch must be explained by
        # the distance to that oracle's separating knot (a few float32 spacings), none is tolerated otherwise.
        from test_gpu_parity import bin_mismatch_ulps
        of.forward(torch.from_numpy(xs).double())
        assert np.array_equal(np.stack([l.last_bins.numpy() for l in reversed(of.layers)]), d["bins"])
        ulps = bin_mismatch_ulps(of, got)
        assert all(u <= 16 for u in ulps), ulps
        assert len(ulps) <= 1e-4 * d["bins"].size, (len(ulps), d["bins"].size)
