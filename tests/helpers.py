"""Shared test helpers: build the same flow in the oracle and in the CUDA package."""
import numpy as np
import torch

import nf_oracle as O

TDT = {np.float32: torch.float32, np.float64: torch.float64}


def oracle_flow(kind, dim, dtype, seed=123, **kw):
    rng = np.random.Generator(np.random.PCG64(seed))
    td = TDT[dtype]
    if kind == "planar":
        return O.planarflow(dim, kw.get("nlayers", 10), td, rng)
    if kind == "radial":
        return O.radialflow(dim, kw.get("nlayers", 10), td, rng)
    if kind == "realnvp":
        return O.realnvp(dim, kw.get("hdims", [32, 32]), kw.get("nlayers", 2), td, rng)
    if kind == "nsf":
        return O.nsf(dim, kw.get("hdims", [32, 32]), kw.get("K", 10), kw.get("B", 5.0), kw.get("nlayers", 2), td, rng)
    if kind == "hamiltonian":   # dim = 2h; LeapFrog on Funnel(h, -8, 5) (example/demo_hamiltonian_flow.jl:104-147), jittered theta
        import math
        tgt = O.Funnel(dim // 2, -8.0, 5.0)
        f = O.hamiltonian_flow(tgt, kw.get("nlayers", 15), kw.get("L", 3), math.log(0.05), dtype=td)
        th = f.theta().double().numpy()
        f.set_theta(torch.from_numpy(th + 0.05 * rng.standard_normal(th.size)).to(td))
        return f
    if kind == "hamiltonian_logreg":   # BASELINE config 5 at its size: dim = 2h, LeapFrog on the h-D logistic-regression posterior
        import math
        tgt = O.synthetic_logreg(dim // 2, kw.get("n_data", 256))
        f = O.hamiltonian_flow(tgt, kw.get("nlayers", 15), kw.get("L", 3), math.log(kw.get("eps", 0.02)), dtype=td)
        th = f.theta().double().numpy()
        f.set_theta(torch.from_numpy(th + 0.05 * rng.standard_normal(th.size)).to(td))
        return f
    raise ValueError(kind)


def gpu_flow(nf, of, dtype, base=None):
    """Mirror an oracle Flow as a package Flow with identical structure and theta."""
    layers = []
    for l in of.layers:
        if isinstance(l, O.Planar):
            layers.append(nf.PlanarLayer(of.dim))
        elif isinstance(l, O.Radial):
            layers.append(nf.RadialLayer(of.dim))
        elif isinstance(l, O.Shift):
            layers.append(nf.Shift(l.a.detach().numpy()))
        elif isinstance(l, O.Scale):
            layers.append(nf.Scale(l.a.detach().numpy()))
        elif isinstance(l, O.MomentumAffine):
            layers.append(nf.MomentumAffine(l.shift.detach().numpy(), l.scale.detach().numpy()))
        elif isinstance(l, O.LeapFrog):
            layers.append(nf.LeapFrog(l.log_eps.numel(), l.log_eps.detach().numpy(), l.L, gpu_target(nf, l.target)))
        elif isinstance(l, O.AffineCoupling):
            hd = [W.shape[1] for W in l.s.Wts[:-1]]
            layers.append(nf.AffineCoupling(of.dim, hd, l.idx1, dtype))
        elif isinstance(l, O.NeuralSplineCoupling):
            hd = [W.shape[1] for W in l.nn.Wts[:-1]]
            layers.append(nf.NeuralSplineCoupling(of.dim, hd, l.K, l.B, l.idx1, dtype))
        else:
            raise TypeError(l)
    mu = np.zeros(of.dim) if of.base_mu is None else of.base_mu.numpy()
    sg = np.ones(of.dim) if of.base_sigma is None else of.base_sigma.numpy()
    if getattr(of, "base_chol", None) is not None:
        Lc = of.base_chol.double().numpy()
        sg = Lc @ Lc.T                                  # MvNormal(mu, Sigma): full covariance
    f = nf.Flow(layers, nf.MvNormal(mu, sg), dtype)
    f.theta = of.theta().numpy().astype(dtype)
    return f


def oracle_target(name, dim):
    if name == "banana":
        return O.Banana(dim, 1.0, 10.0)
    if name == "funnel":
        return O.Funnel(dim, 0.0, 9.0)
    if name == "warped":
        return O.WarpedGauss(1.0, 0.12)
    if name == "cross":
        return O.Cross(dim, 2.0, 0.15)
    if name == "joint_funnel":
        return O.JointTarget(O.Funnel(dim // 2, -8.0, 5.0))
    if name == "joint_logreg":
        return O.JointTarget(O.synthetic_logreg(dim // 2, 256))
    if name == "diag":
        rng = np.random.Generator(np.random.PCG64(7))
        return O.DiagNormal(rng.standard_normal(dim), rng.uniform(0.5, 1.5, dim))
    raise ValueError(name)


def gpu_target(nf, ot):
    if isinstance(ot, O.Banana):
        return nf.Banana(ot.dim, ot.b, ot.var)
    if isinstance(ot, O.Funnel):
        return nf.Funnel(ot.dim, ot.mu, ot.sigma)
    if isinstance(ot, O.WarpedGauss):
        return nf.WarpedGauss(ot.s1, ot.s2)
    if isinstance(ot, O.Cross):
        return nf.Cross(ot.mu, ot.sigma, ot.dim)
    if isinstance(ot, O.DiagNormal):
        return nf.DiagNormal(ot.mu.numpy(), ot.sigma.numpy())
    if isinstance(ot, O.LogReg):
        return nf.LogReg(ot.X.numpy(), ot.y.numpy(), ot.sigma0)
    if isinstance(ot, O.JointTarget):
        return nf.JointTarget(gpu_target(nf, ot.inner))
    raise TypeError(ot)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def z0(n, dim, dtype, seed=2024):
    return O.synthetic_z0(n, dim, seed).astype(dtype)
