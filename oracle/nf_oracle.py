"""CPU oracle for the NormalizingFlows.jl training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (PyTorch CPU tensors + autograd) of the algorithm the reference
runs for `elbo` / `elbo_batch` / `loglikelihood` value+gradient.  It exists so the CUDA path can be
checked; it is NOT part of the product.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

PARITY STATUS: **parity unpinned** for fixed-value outputs.  Neither Julia nor the third-party
packages that hold most of the arithmetic (Bijectors.jl 0.15/0.16, MonotonicSplines.jl 0.3.3,
Flux 0.16 / NNlib, Optimisers 0.4 / Functors 0.5, Distributions 0.25) exist in this image, and the
reference's tests carry no golden vectors.  The oracle is therefore pinned
only through the reference's own *property* tests, which `tests/test_oracle.py` ports:
  test/objectives.jl:14-26  analytic ELBO == 0 for an exact Shift∘Scale flow, logpdf(flow,x)+el == logp(x)
  test/flow.jl:25-39,92-106,158-172,224-238   inverse consistency (rtol 1e-6 / 1e-4), odd d = 5
  test/flow.jl:42-61        finite elbo / elbo_batch at n = 64 and n = 1
  test/interface.jl:28-50   Shift∘Scale training converges to (10,10,2,2)
plus autograd-vs-finite-difference checks and an independent pin of every hand-written log|det J| (planar, radial, affine
and spline couplings, leapfrog) against log|det| of the autograd Jacobian of the forward map.  `julia/crosscheck.jl` (shipped, unexecuted here) turns
the UNVERIFIED switches below into pass/fail on a machine that has Julia.

Layout convention: a Julia `d×N` column-major batch is a row-major `[N, d]` tensor here (each
sample's d numbers contiguous).  theta follows `Optimisers.destructure` order.

Reference lines followed (relative to /root/reference):
  src/objectives/elbo.jl:4-7        elbo_single_sample
  src/objectives/elbo.jl:31-34      elbo   (map over columns, mean)
  src/objectives/elbo.jl:65-70,89-92  _batched_elbos / elbo_batch
  src/objectives/loglikelihood.jl:26-33  loglikelihood
  src/flows/realnvp.jl:42-110,132-145,170-180   AffineCoupling fwd/inv, RealNVP_layer, realnvp
  src/flows/neuralspline.jl:44-71,94-140,169-184,218-234  NeuralSplineCoupling, NSF_layer, nsf
  src/flows/utils.jl:23-26,71-100   create_flow, fnn
  src/flows/planar_radial.jl:21-29,52-60  planarflow, radialflow
  test/ext/CUDA/cuda.jl:12-30       planar get_u_hat / _transform restated in-tree
  example/targets/*.jl              Banana, Funnel, WarpedGauss, Cross log-densities
Third-party semantics: SURVEY.md Appendix A (A.1 planar, A.2 radial, A.3 PartitionMask,
A.4 MonotonicSplines RQS, A.5 composition, A.6 Flux Dense, A.7 destructure, A.8 MvNormal).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

LOG2PI = math.log(2.0 * math.pi)

# ----------------------------------------------------------------------------------------------
# UNVERIFIED upstream details, each behind one switch (SURVEY App. A "UNVERIFIED" list).
# ----------------------------------------------------------------------------------------------
#: MonotonicSplines.rqs_params_from_nn reshape: "block" = (3K-1, c, N) i.e. each transformed
#: coordinate owns a contiguous block of 3K-1 conditioner outputs [K widths | K heights | K-1 derivs];
#: "interleaved" = (c, 3K-1, N) coordinate-fastest.
RQS_PARAM_LAYOUT = "block"
#: `searchsortedfirst(pX, x) - 1`: bins are (pX[k], pX[k+1]]  ("right-closed"); the alternative is
#: `searchsortedlast` = [pX[k], pX[k+1]).
RQS_BIN_RIGHT_CLOSED = True
#: softmax without max-subtraction (exp.(x) ./ sum(exp.(x))) as recalled from MonotonicSplines._softmax
RQS_SOFTMAX_SUBTRACT_MAX = False


def _softplus(x: torch.Tensor) -> torch.Tensor:
    """LogExpFunctions.log1pexp (Bijectors planar/radial): stable softplus."""
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-torch.abs(x)))


def _naive_softplus(x: torch.Tensor) -> torch.Tensor:
    """MonotonicSplines._softplus(x) = log(exp(x) + 1)  (App. A.4)."""
    return torch.log(torch.exp(x) + 1)


def leakyrelu(x: torch.Tensor) -> torch.Tensor:
    """NNlib.leakyrelu, slope 0.01 (App. A.6)."""
    return torch.where(x > 0, x, 0.01 * x)


# ----------------------------------------------------------------------------------------------
# Layers.  Every layer exposes: params() -> list of tensors in destructure order,
# forward(x[N,d]) -> (y[N,d], logdet[N]) and inverse(y) -> (x, logdet_inv).
# ----------------------------------------------------------------------------------------------
class Layer:
    kind = "abstract"

    def params(self) -> List[torch.Tensor]:
        raise NotImplementedError

    def forward(self, x):
        raise NotImplementedError

    def inverse(self, y):
        raise NotImplementedError

    def n_params(self) -> int:
        return sum(p.numel() for p in self.params())


class Shift(Layer):
    """Bijectors.Shift(a): y = x + a, logdet 0 (used by test/objectives.jl:9, test/interface.jl:22-24)."""
    kind = "shift"

    def __init__(self, a: torch.Tensor):
        self.a = a

    def params(self):
        return [self.a]

    def forward(self, x):
        return x + self.a, torch.zeros(x.shape[0], dtype=x.dtype)

    def inverse(self, y):
        return y - self.a, torch.zeros(y.shape[0], dtype=y.dtype)


class Scale(Layer):
    """Bijectors.Scale(a) elementwise: y = a .* x, logdet = sum(log|a|)."""
    kind = "scale"

    def __init__(self, a: torch.Tensor):
        self.a = a

    def params(self):
        return [self.a]

    def forward(self, x):
        ld = torch.log(torch.abs(self.a)).sum()
        return x * self.a, ld.expand(x.shape[0])

    def inverse(self, y):
        ld = -torch.log(torch.abs(self.a)).sum()
        return y / self.a, ld.expand(y.shape[0])


class Planar(Layer):
    """Bijectors.PlanarLayer (App. A.1; in-tree restatement test/ext/CUDA/cuda.jl:12-30)."""
    kind = "planar"

    def __init__(self, w, u, b):
        self.w, self.u, self.b = w, u, b

    def params(self):
        return [self.w, self.u, self.b]

    def u_hat(self):
        s = torch.dot(self.w, self.u)
        u_hat = self.u + ((_softplus(-s) - 1) / torch.sum(self.w * self.w)) * self.w
        wtu = _softplus(s) - 1
        return u_hat, wtu

    def forward(self, x):
        u_hat, wtu = self.u_hat()
        a = x @ self.w + self.b[0]
        t = torch.tanh(a)
        y = x + t[:, None] * u_hat[None, :]
        ld = torch.log1p(wtu * (1 - t * t))
        return y, ld

    def inverse(self, y, iters: int = 200):
        # solve alpha + wtu*tanh(alpha + b) = w.y by bisection on [wy-|wtu|, wy+|wtu|] (App. A.1)
        u_hat, wtu = self.u_hat()
        wy = y @ self.w
        b = self.b[0]
        with torch.no_grad():
            lo = wy - torch.abs(wtu)
            hi = wy + torch.abs(wtu)
            for _ in range(iters):
                mid = 0.5 * (lo + hi)
                f = mid + wtu * torch.tanh(mid + b) - wy
                lo = torch.where(f < 0, mid, lo)
                hi = torch.where(f < 0, hi, mid)
            alpha = 0.5 * (lo + hi)
        # one Newton polish step that is differentiable w.r.t. parameters (implicit function)
        f = alpha + wtu * torch.tanh(alpha + b) - wy
        fp = 1 + wtu * (1 - torch.tanh(alpha + b) ** 2)
        alpha = alpha - f / fp
        t = torch.tanh(alpha + b)
        x = y - t[:, None] * u_hat[None, :]
        ld = -torch.log1p(wtu * (1 - t * t))
        return x, ld


class Radial(Layer):
    """Bijectors.RadialLayer (App. A.2)."""
    kind = "radial"

    def __init__(self, alpha_, beta, z0):
        self.alpha_, self.beta, self.z0 = alpha_, beta, z0

    def params(self):
        return [self.alpha_, self.beta, self.z0]

    def _ab(self):
        alpha = _softplus(self.alpha_[0])
        beta_hat = -alpha + _softplus(self.beta[0])
        return alpha, beta_hat

    def forward(self, x):
        d = x.shape[1]
        alpha, bh = self._ab()
        diff = x - self.z0
        r = torch.sqrt(torch.sum(diff * diff, dim=1))
        h = 1 / (alpha + r)
        y = x + (bh * h)[:, None] * diff
        ld = (d - 1) * torch.log1p(bh * h) + torch.log1p(bh * h - bh * h * h * r)
        return y, ld

    def inverse(self, y):
        d = y.shape[1]
        alpha, bh = self._ab()
        diff = y - self.z0
        rho = torch.sqrt(torch.sum(diff * diff, dim=1))
        a = (alpha + bh) - rho
        r = (torch.sqrt(a * a + 4 * alpha * rho) - a) / 2
        x = self.z0 + ((alpha + r) / (alpha + bh + r))[:, None] * diff
        h = 1 / (alpha + r)
        ld = -((d - 1) * torch.log1p(bh * h) + torch.log1p(bh * h - bh * h * h * r))
        return x, ld


class MLP:
    """`fnn` (src/flows/utils.jl:71-100): Dense(in,h1,leakyrelu) ... Dense(h_end,out[,act]).

    Weights are stored as the Julia `out×in` matrix *in Julia memory order*, i.e. as a row-major
    [in, out] tensor `Wt` (= transpose), so that `Wt.reshape(-1)` is exactly `vec(W)`.
    """

    def __init__(self, Wts: List[torch.Tensor], bs: List[torch.Tensor], out_act: Optional[str]):
        self.Wts, self.bs, self.out_act = Wts, bs, out_act

    def params(self):
        out = []
        for Wt, b in zip(self.Wts, self.bs):
            out += [Wt, b]
        return out

    def __call__(self, x):
        h = x
        n = len(self.Wts)
        for i, (Wt, b) in enumerate(zip(self.Wts, self.bs)):
            h = h @ Wt + b
            if i < n - 1:
                h = leakyrelu(h)
            elif self.out_act == "tanh":
                h = torch.tanh(h)
        return h


def partition_indices(dim: int, mask_idx: Sequence[int]) -> Tuple[List[int], List[int]]:
    """Bijectors.PartitionMask(dim, idx) (App. A.3), 0-based: (transformed idx, sorted complement)."""
    idx = list(mask_idx)
    s = set(idx)
    comp = [i for i in range(dim) if i not in s]
    return idx, comp


class AffineCoupling(Layer):
    """src/flows/realnvp.jl:33-110."""
    kind = "affine_coupling"

    def __init__(self, dim: int, mask_idx: Sequence[int], s: MLP, t: MLP):
        self.dim = dim
        self.idx1, self.idx2 = partition_indices(dim, mask_idx)
        self.s, self.t = s, t

    def params(self):
        return self.s.params() + self.t.params()

    def forward(self, x):                               # realnvp.jl:77-83
        x1, x2 = x[:, self.idx1], x[:, self.idx2]
        s = self.s(x2)
        y1 = torch.exp(s) * x1 + self.t(x2)
        y = x.clone()
        y[:, self.idx1] = y1
        return y, s.sum(dim=1)

    def inverse(self, y):                               # realnvp.jl:99-110
        y1, y2 = y[:, self.idx1], y[:, self.idx2]
        s = self.s(y2)
        x1 = (y1 - self.t(y2)) * torch.exp(-s)
        x = y.clone()
        x[:, self.idx1] = x1
        return x, -s.sum(dim=1)


class MomentumAffine(Layer):
    """`Stacked((identity, Shift(b) ∘ Scale(a)), [1:d, d+1:2d])` -- the momentum normalisation layer of
    example/demo_hamiltonian_flow.jl:94-99: position block untouched, rho <- a .* rho .+ b.  theta: shift(d), scale(d)."""
    kind = "momentum_affine"

    def __init__(self, shift, scale):
        self.shift, self.scale = shift, scale

    def params(self):
        return [self.shift, self.scale]

    def forward(self, z):
        d = self.shift.numel()
        y = torch.cat([z[:, :d], z[:, d:] * self.scale + self.shift], dim=1)
        return y, torch.log(torch.abs(self.scale)).sum().expand(z.shape[0])

    def inverse(self, y):
        d = self.shift.numel()
        z = torch.cat([y[:, :d], (y[:, d:] - self.shift) / self.scale], dim=1)
        return z, (-torch.log(torch.abs(self.scale)).sum()).expand(y.shape[0])


class LeapFrog(Layer):
    """`LeapFrog` bijector of example/demo_hamiltonian_flow.jl:27-91: L leapfrog steps on z = [x, rho] with per-dimension
    step sizes exp(log_eps) (the only trainable field, `@functor LeapFrog (logϵ,)`, :39) and the target score inside the
    transform (:49-61).  Symplectic: logabsdetjac = 0 (:84-91).  `score` must be differentiable (second-order terms)."""
    kind = "leapfrog"

    def __init__(self, log_eps, L, score, target=None):
        self.log_eps, self.L, self.score, self.target = log_eps, int(L), score, target

    def params(self):
        return [self.log_eps]

    def _run(self, z, eps):
        d = self.log_eps.numel()
        x, v = z[:, :d], z[:, d:]
        v = v + eps / 2 * self.score(x)
        for _ in range(self.L - 1):
            x = x + eps * v
            v = v + eps * self.score(x)
        x = x + eps * v
        v = v + eps / 2 * self.score(x)
        return torch.cat([x, v], dim=1)

    def forward(self, z):
        return self._run(z, torch.exp(self.log_eps)), torch.zeros(z.shape[0], dtype=z.dtype)

    def inverse(self, z):
        return self._run(z, -torch.exp(self.log_eps)), torch.zeros(z.shape[0], dtype=z.dtype)


# ---- rational-quadratic splines (MonotonicSplines v0.3.3, App. A.4) --------------------------
def rqs_slot_rows(c: int, K: int):
    """Row -> (coordinate, slot) map of the (3K-1)*c conditioner outputs.  Returns three index
    arrays of shape [c, K], [c, K], [c, K-1] giving the rows of width / height / derivative logits."""
    P = 3 * K - 1
    if RQS_PARAM_LAYOUT == "block":
        base = np.arange(c)[:, None] * P
        w = base + np.arange(K)[None, :]
        h = base + K + np.arange(K)[None, :]
        dv = base + 2 * K + np.arange(K - 1)[None, :]
    else:  # coordinate-fastest
        w = np.arange(K)[None, :] * c + np.arange(c)[:, None]
        h = (K + np.arange(K))[None, :] * c + np.arange(c)[:, None]
        dv = (2 * K + np.arange(K - 1))[None, :] * c + np.arange(c)[:, None]
    return w, h, dv


def _seq_cumsum(p: torch.Tensor) -> torch.Tensor:
    """Canonical left-to-right cumulative sum in the working dtype (SURVEY section 7 'bit-exact bins')."""
    out = []
    acc = torch.zeros_like(p[..., 0])
    for k in range(p.shape[-1]):
        acc = acc + p[..., k]
        out.append(acc)
    return torch.stack(out, dim=-1)


def rqs_params_from_nn(theta_raw: torch.Tensor, c: int, B: float):
    """theta_raw [N, (3K-1)c] -> pX, pY [N, c, K+1], dYdX [N, c, K+1]."""
    N = theta_raw.shape[0]
    K = (theta_raw.shape[1] // c + 1) // 3
    wr, hr, dr = rqs_slot_rows(c, K)
    tw = theta_raw[:, torch.as_tensor(wr)]              # [N, c, K]
    th = theta_raw[:, torch.as_tensor(hr)]
    td = theta_raw[:, torch.as_tensor(dr)]              # [N, c, K-1]

    def knots(t):
        if RQS_SOFTMAX_SUBTRACT_MAX:
            t = t - t.max(dim=-1, keepdim=True).values
        e = torch.exp(t)
        ssum = torch.zeros_like(e[..., 0])
        for k in range(K):                              # sequential sum, left to right
            ssum = ssum + e[..., k]
        p = e / ssum[..., None]
        cs = _seq_cumsum(p)
        kn = (2 * B) * cs - B
        left = torch.full_like(kn[..., :1], -B)
        return torch.cat([left, kn], dim=-1)

    pX, pY = knots(tw), knots(th)
    one = torch.ones_like(td[..., :1])
    dYdX = torch.cat([one, _naive_softplus(td), one], dim=-1)
    return pX, pY, dYdX


def rqs_bin_index(knots: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """k = searchsortedfirst(knots, v) - 1 (number of knots strictly below v), 1-based bins 1..K are inside."""
    if RQS_BIN_RIGHT_CLOSED:
        return (knots < v[..., None]).sum(dim=-1)
    return (knots <= v[..., None]).sum(dim=-1)


def _gather(t, k):
    return torch.gather(t, -1, k[..., None]).squeeze(-1)


def rqs_forward(x: torch.Tensor, pX, pY, dYdX):
    """x [N, c] -> y [N, c], logJ [N, c] (summed over c by the caller), bins [N, c] (int64)."""
    K = pX.shape[-1] - 1
    k = rqs_bin_index(pX, x)
    inside = (k >= 1) & (k <= K)
    kk = torch.clamp(k, 1, K)
    x0, x1 = _gather(pX, kk - 1), _gather(pX, kk)
    y0, y1 = _gather(pY, kk - 1), _gather(pY, kk)
    d0, d1 = _gather(dYdX, kk - 1), _gather(dYdX, kk)
    dx, dy = x1 - x0, y1 - y0
    s = dy / dx
    xi = (x - x0) / dx
    om = 1 - xi
    den = s + (d1 + d0 - 2 * s) * xi * om
    y_in = y0 + dy * (s * xi * xi + d0 * xi * om) / den
    lj_in = torch.log(torch.abs(s * s * (d1 * xi * xi + 2 * s * xi * om + d0 * om * om))) - 2 * torch.log(torch.abs(den))
    y = torch.where(inside, y_in, x)
    lj = torch.where(inside, lj_in, torch.zeros_like(lj_in))
    return y, lj, k


def rqs_inverse(y: torch.Tensor, pX, pY, dYdX):
    K = pX.shape[-1] - 1
    k = rqs_bin_index(pY, y)
    inside = (k >= 1) & (k <= K)
    kk = torch.clamp(k, 1, K)
    x0, x1 = _gather(pX, kk - 1), _gather(pX, kk)
    y0, y1 = _gather(pY, kk - 1), _gather(pY, kk)
    d0, d1 = _gather(dYdX, kk - 1), _gather(dYdX, kk)
    dx, dy = x1 - x0, y1 - y0
    s = dy / dx
    yr = y - y0
    t = d1 + d0 - 2 * s
    a = dy * (s - d0) + yr * t
    b = dy * d0 - yr * t
    cq = -s * yr
    xi = 2 * cq / (-b - torch.sqrt(b * b - 4 * a * cq))
    om = 1 - xi
    den = s + t * xi * om
    x_in = xi * dx + x0
    lj_in = -(torch.log(torch.abs(s * s * (d1 * xi * xi + 2 * s * xi * om + d0 * om * om))) - 2 * torch.log(torch.abs(den)))
    x = torch.where(inside, x_in, y)
    lj = torch.where(inside, lj_in, torch.zeros_like(lj_in))
    return x, lj, k


class NeuralSplineCoupling(Layer):
    """src/flows/neuralspline.jl:35-140."""
    kind = "spline_coupling"

    def __init__(self, dim: int, K: int, B: float, mask_idx: Sequence[int], nn: MLP):
        self.dim, self.K, self.B = dim, K, float(B)
        self.idx1, self.idx2 = partition_indices(dim, mask_idx)
        self.nn = nn
        self.last_bins = None

    def params(self):
        return self.nn.params()

    def _spline_params(self, x2):                       # neuralspline.jl:65-71
        return rqs_params_from_nn(self.nn(x2), len(self.idx1), self.B)

    def forward(self, x):                               # neuralspline.jl:102-108
        x1, x2 = x[:, self.idx1], x[:, self.idx2]
        pX, pY, dYdX = self._spline_params(x2)
        y1, lj, k = rqs_forward(x1, pX, pY, dYdX)
        self.last_bins = k.detach()
        self.last_knots, self.last_v = pX.detach(), x1.detach()   # searched knots / searched values (bin-parity tests)
        y = x.clone()
        y[:, self.idx1] = y1
        return y, lj.sum(dim=1)

    def inverse(self, y):                               # neuralspline.jl:133-140
        y1, y2 = y[:, self.idx1], y[:, self.idx2]
        pX, pY, dYdX = self._spline_params(y2)
        x1, lj, k = rqs_inverse(y1, pX, pY, dYdX)
        self.last_bins = k.detach()
        self.last_knots, self.last_v = pY.detach(), y1.detach()
        x = y.clone()
        x[:, self.idx1] = x1
        return x, lj.sum(dim=1)


# ----------------------------------------------------------------------------------------------
# Flow = TransformedDistribution(q0 = diag MvNormal, reduce(∘, Ls))
# ----------------------------------------------------------------------------------------------
@dataclass
class Flow:
    """`create_flow(Ls, q0)` (src/flows/utils.jl:23-26).  `layers` is `Ls` in theta order; the
    transform applies `Ls[end]` first (App. A.5).  Base q0 = MvNormal(mu, Diagonal(sigma.^2))."""
    dim: int
    layers: List[Layer]
    base_mu: Optional[torch.Tensor] = None
    base_sigma: Optional[torch.Tensor] = None
    base_chol: Optional[torch.Tensor] = None            # full covariance Sigma = L L^T (lower L); overrides base_sigma
    dtype: torch.dtype = torch.float64

    def params(self) -> List[torch.Tensor]:
        out = []
        for l in self.layers:
            out += l.params()
        return out

    def n_params(self) -> int:
        return sum(p.numel() for p in self.params())

    def theta(self) -> torch.Tensor:
        return torch.cat([p.detach().reshape(-1) for p in self.params()]).clone()

    def set_theta(self, theta: torch.Tensor, requires_grad: bool = False):
        """`re(theta)` of Optimisers.destructure: rebinds every leaf to a view of `theta`."""
        theta = theta.to(self.dtype)
        if requires_grad:
            theta = theta.clone().requires_grad_(True)
        off = 0
        for l in self.layers:
            _rebind(l, theta, off)
            off += l.n_params()
        assert off == theta.numel(), (off, theta.numel())
        return theta

    def base_logpdf(self, x):                           # App. A.8
        mu = self.base_mu if self.base_mu is not None else torch.zeros(self.dim, dtype=x.dtype)
        if self.base_chol is not None:
            # Distributions.logpdf(MvNormal(mu, Sigma)): -(d log 2pi + logdet Sigma)/2 - |L^-1 (x - mu)|^2 / 2
            L = self.base_chol.to(x.dtype)
            z = torch.linalg.solve_triangular(L, (x - mu).T, upper=False).T
            return -0.5 * self.dim * LOG2PI - torch.log(torch.diagonal(L)).sum() - 0.5 * (z * z).sum(dim=1)
        sg = self.base_sigma if self.base_sigma is not None else torch.ones(self.dim, dtype=x.dtype)
        z = (x - mu) / sg
        return -0.5 * self.dim * LOG2PI - torch.log(sg).sum() - 0.5 * (z * z).sum(dim=1)

    def base_sample(self, z_std):
        """MvNormal sampling as ext/NormalizingFlowsCUDAExt.jl:43-48: randn -> unwhiten -> + mu."""
        mu = self.base_mu if self.base_mu is not None else torch.zeros(self.dim, dtype=z_std.dtype)
        if self.base_chol is not None:                  # unwhiten!(Sigma, x) .+ mu (ext/NormalizingFlowsCUDAExt.jl:45-46)
            return z_std @ self.base_chol.to(z_std.dtype).T + mu
        sg = self.base_sigma if self.base_sigma is not None else torch.ones(self.dim, dtype=z_std.dtype)
        return z_std * sg + mu

    def forward(self, x):
        ld = torch.zeros(x.shape[0], dtype=x.dtype)
        for l in reversed(self.layers):
            x, l_ld = l.forward(x)
            ld = ld + l_ld
        return x, ld

    def inverse(self, y):
        ld = torch.zeros(y.shape[0], dtype=y.dtype)
        for l in self.layers:
            y, l_ld = l.inverse(y)
            ld = ld + l_ld
        return y, ld

    def logpdf(self, y):                                # Bijectors logpdf(td, y), App. A.5
        x, ld = self.inverse(y)
        return self.base_logpdf(x) + ld


def _rebind(layer: Layer, theta: torch.Tensor, off: int):
    def take(shape):
        nonlocal off
        n = int(np.prod(shape))
        v = theta[off:off + n].reshape(shape)
        off += n
        return v

    if isinstance(layer, (Shift, Scale)):
        layer.a = take(layer.a.shape)
    elif isinstance(layer, MomentumAffine):
        layer.shift = take(layer.shift.shape); layer.scale = take(layer.scale.shape)
    elif isinstance(layer, LeapFrog):
        layer.log_eps = take(layer.log_eps.shape)
    elif isinstance(layer, Planar):
        layer.w = take(layer.w.shape); layer.u = take(layer.u.shape); layer.b = take(layer.b.shape)
    elif isinstance(layer, Radial):
        layer.alpha_ = take(layer.alpha_.shape); layer.beta = take(layer.beta.shape); layer.z0 = take(layer.z0.shape)
    elif isinstance(layer, AffineCoupling):
        for m in (layer.s, layer.t):
            for i in range(len(m.Wts)):
                m.Wts[i] = take(m.Wts[i].shape); m.bs[i] = take(m.bs[i].shape)
    elif isinstance(layer, NeuralSplineCoupling):
        m = layer.nn
        for i in range(len(m.Wts)):
            m.Wts[i] = take(m.Wts[i].shape); m.bs[i] = take(m.bs[i].shape)
    else:
        raise TypeError(layer)


# ----------------------------------------------------------------------------------------------
# Constructors mirroring the reference API (synthetic initialisation of SURVEY section 8d)
# ----------------------------------------------------------------------------------------------
def _glorot(rng: np.random.Generator, n_in: int, n_out: int, dtype) -> torch.Tensor:
    """Flux.glorot_uniform for Dense(in,out): U(+-sqrt(6/(in+out))), drawn in Float32 then cast
    (App. A.6).  Returned in Julia memory order ([in, out] row-major == vec(W) of out×in)."""
    lim = math.sqrt(6.0 / (n_in + n_out))
    W = rng.uniform(-lim, lim, size=(n_in, n_out)).astype(np.float32)
    return torch.from_numpy(W).to(dtype)


def fnn(rng, n_in: int, hdims: Sequence[int], n_out: int, out_act: Optional[str], dtype) -> MLP:
    dims = [n_in] + list(hdims) + [n_out]
    Wts, bs = [], []
    for a, b in zip(dims[:-1], dims[1:]):
        Wts.append(_glorot(rng, a, b, dtype))
        bs.append(torch.zeros(b, dtype=dtype))
    return MLP(Wts, bs, out_act)


def _randn(rng, n, dtype):
    return torch.from_numpy(rng.standard_normal(n)).to(dtype)


def planarflow(dim: int, nlayers: int, dtype=torch.float64, rng=None) -> Flow:
    """src/flows/planar_radial.jl:21-29; parameters ~ randn (Bijectors ctor, App. A.1)."""
    rng = rng or np.random.Generator(np.random.PCG64(123))
    Ls = [Planar(_randn(rng, dim, dtype), _randn(rng, dim, dtype), _randn(rng, 1, dtype)) for _ in range(nlayers)]
    return Flow(dim, Ls, dtype=dtype)


def radialflow(dim: int, nlayers: int, dtype=torch.float64, rng=None) -> Flow:
    """src/flows/planar_radial.jl:52-60."""
    rng = rng or np.random.Generator(np.random.PCG64(123))
    Ls = [Radial(_randn(rng, 1, dtype), _randn(rng, 1, dtype), _randn(rng, dim, dtype)) for _ in range(nlayers)]
    return Flow(dim, Ls, dtype=dtype)


def realnvp(dim: int, hdims: Sequence[int], nlayers: int, dtype=torch.float64, rng=None) -> Flow:
    """src/flows/realnvp.jl:132-145,170-180: each RealNVP_layer = af(1:2:d) ∘ af(2:2:d)."""
    rng = rng or np.random.Generator(np.random.PCG64(123))
    Ls: List[Layer] = []
    for _ in range(nlayers):
        for mask in (list(range(0, dim, 2)), list(range(1, dim, 2))):
            c = len(mask)
            s = fnn(rng, dim - c, hdims, c, "tanh", dtype)
            t = fnn(rng, dim - c, hdims, c, None, dtype)
            Ls.append(AffineCoupling(dim, mask, s, t))
    return Flow(dim, Ls, dtype=dtype)


def nsf(dim: int, hdims: Sequence[int], K: int, B: float, nlayers: int, dtype=torch.float64, rng=None) -> Flow:
    """src/flows/neuralspline.jl:169-184,218-230."""
    rng = rng or np.random.Generator(np.random.PCG64(123))
    Ls: List[Layer] = []
    for _ in range(nlayers):
        for mask in (list(range(0, dim, 2)), list(range(1, dim, 2))):
            c = len(mask)
            nn = fnn(rng, dim - c, hdims, (3 * K - 1) * c, None, dtype)
            Ls.append(NeuralSplineCoupling(dim, K, B, mask, nn))
    return Flow(dim, Ls, dtype=dtype)


def shift_scale_flow(shift, scale, dtype=torch.float64, base_sigma=None) -> Flow:
    """Bijectors.Shift(mu) ∘ Bijectors.Scale(s) (test/objectives.jl:9): theta = [shift; scale]."""
    shift = torch.as_tensor(shift, dtype=dtype)
    scale = torch.as_tensor(scale, dtype=dtype)
    return Flow(shift.numel(), [Shift(shift), Scale(scale)], dtype=dtype,
                base_sigma=None if base_sigma is None else torch.as_tensor(base_sigma, dtype=dtype))


# ----------------------------------------------------------------------------------------------
# Targets (example/targets/*.jl; formulas SURVEY App. B)
# ----------------------------------------------------------------------------------------------
class Target:
    kind = "abstract"
    params: Tuple[float, ...] = ()

    def logp(self, y):
        raise NotImplementedError


class Banana(Target):
    """example/targets/banana.jl:77-83."""
    kind = "banana"

    def __init__(self, dim, b, var):
        self.dim, self.b, self.var = dim, float(b), float(var)
        self.params = (self.b, self.var)

    def logp(self, y):
        d, b, v = self.dim, self.b, self.var
        u1 = y[:, 0]
        u2 = y[:, 1] + b * u1 * u1 - v * b
        q = u1 * u1 / v + u2 * u2
        if d > 2:
            q = q + (y[:, 2:] ** 2).sum(dim=1)
        logz = (math.log(v) / d + LOG2PI) * d / 2
        return -logz - q / 2


class Funnel(Target):
    """example/targets/neal_funnel.jl:54-61."""
    kind = "funnel"

    def __init__(self, dim, mu=0.0, sigma=9.0):
        self.dim, self.mu, self.sigma = dim, float(mu), float(sigma)
        self.params = (self.mu, self.sigma)

    def logp(self, y):
        d = self.dim
        x1 = y[:, 0]
        lp1 = -0.5 * LOG2PI - math.log(self.sigma) - (x1 - self.mu) ** 2 / (2 * self.sigma ** 2)
        ss = (y[:, 1:] ** 2).sum(dim=1)
        lp2 = -(d - 1) / 2 * (LOG2PI + x1) - 0.5 * torch.exp(-x1) * ss
        return lp1 + lp2


class WarpedGauss(Target):
    """example/targets/warped_gaussian.jl:54-68,81-87 (the +log r term is as coded)."""
    kind = "warped_gauss"

    def __init__(self, s1=1.0, s2=0.12):
        self.dim, self.s1, self.s2 = 2, float(s1), float(s2)
        self.params = (self.s1, self.s2)

    def logp(self, y):
        x, yy = y[:, 0], y[:, 1]
        r = torch.sqrt(x * x + yy * yy)
        th = torch.atan2(yy, x) + r / 2
        z1, z2 = r * torch.cos(th), r * torch.sin(th)
        return (-(z1 * z1 / self.s1 ** 2 + z2 * z2 / self.s2 ** 2) / 2 - LOG2PI
                - math.log(self.s1) - math.log(self.s2) + torch.log(r))


def _cross_block_logp(y2, mu, sigma):
    """example/targets/cross.jl:30-38: means exactly as coded (second coordinate 1 for comps 2,3)."""
    means = [(0.0, mu), (-mu, 1.0), (mu, 1.0), (0.0, -mu)]
    sds = [(sigma, 1.0), (1.0, sigma), (1.0, sigma), (sigma, 1.0)]
    comps = []
    for (m1, m2), (s1, s2) in zip(means, sds):
        lp = (-LOG2PI - math.log(s1) - math.log(s2)
              - 0.5 * (((y2[:, 0] - m1) / s1) ** 2 + ((y2[:, 1] - m2) / s2) ** 2))
        comps.append(lp + math.log(0.25))
    return torch.logsumexp(torch.stack(comps, dim=0), dim=0)


class Cross(Target):
    """2-D Cross, or for dim = 2m the product of m independent Cross blocks (synthetic extension,
    SURVEY section 8d 'Configs made concrete' C4)."""
    kind = "cross"

    def __init__(self, dim=2, mu=2.0, sigma=0.15):
        assert dim % 2 == 0
        self.dim, self.mu, self.sigma = dim, float(mu), float(sigma)
        self.params = (self.mu, self.sigma)

    def logp(self, y):
        out = 0
        for j in range(0, self.dim, 2):
            out = out + _cross_block_logp(y[:, j:j + 2], self.mu, self.sigma)
        return out

    def sample(self, n, rng: np.random.Generator, dtype=torch.float64):
        means = np.array([(0.0, self.mu), (-self.mu, 1.0), (self.mu, 1.0), (0.0, -self.mu)])
        sds = np.array([(self.sigma, 1.0), (1.0, self.sigma), (1.0, self.sigma), (self.sigma, 1.0)])
        out = np.empty((n, self.dim))
        for j in range(0, self.dim, 2):
            k = rng.integers(0, 4, size=n)
            out[:, j:j + 2] = means[k] + sds[k] * rng.standard_normal((n, 2))
        return torch.from_numpy(out).to(dtype)


def _score_funnel(x, mu, sigma):
    """`score(::Funnel, x)` of example/targets/neal_funnel.jl:63-72, batched."""
    d = x.shape[1]
    a = torch.exp(-x[:, :1])
    ss = (x[:, 1:] ** 2).sum(dim=1, keepdim=True)
    g1 = (mu - x[:, :1]) / sigma ** 2 - (d - 1) / 2 + a * ss / 2
    return torch.cat([g1, -a * x[:, 1:]], dim=1)


def _score_banana(x, b, var):
    u2 = x[:, 1:2] + b * x[:, :1] ** 2 - var * b
    g1 = -x[:, :1] / var - 2 * b * x[:, :1] * u2
    return torch.cat([g1, -u2, -x[:, 2:]], dim=1)


def target_score(target):
    """Differentiable score function d logp / dx of a built-in target (used inside LeapFrog)."""
    if isinstance(target, Funnel):
        return lambda x: _score_funnel(x, target.mu, target.sigma)
    if isinstance(target, Banana):
        return lambda x: _score_banana(x, target.b, target.var)
    if isinstance(target, DiagNormal):
        return lambda x: -(x - target.mu.to(x.dtype)) / target.sigma.to(x.dtype) ** 2
    if isinstance(target, LogReg):
        return target.score
    raise TypeError("no score function for %r" % (target,))


class JointTarget(Target):
    """logp_joint(z) = logp(x) + sum(logpdf(Normal(), rho)) on z = [x, rho] (example/demo_hamiltonian_flow.jl:117-124)."""
    kind = "joint"

    def __init__(self, inner):
        self.inner, self.dim = inner, 2 * inner.dim

    def logp(self, z):
        d = self.inner.dim
        rho = z[:, d:]
        return self.inner.logp(z[:, :d]) - 0.5 * d * LOG2PI - 0.5 * (rho * rho).sum(dim=1)


def hamiltonian_flow(target, nlayers: int, L: int, log_eps0: float, dtype=torch.float64) -> "Flow":
    """The flow of example/demo_hamiltonian_flow.jl:128-147: Ls = [momentum_normalization ∘ LeapFrog] x nlayers on top of
    q0 = transformed(MvNormal(0, I_2d), Shift ∘ Scale); `transformed(q0, ts)` composes, so theta ends with q0's Shift, Scale
    (App. A.5) and those are applied first."""
    d = target.dim
    sc = target_score(target)
    Ls: List[Layer] = []
    for _ in range(nlayers):
        Ls.append(MomentumAffine(torch.zeros(d, dtype=dtype), torch.ones(d, dtype=dtype)))
        Ls.append(LeapFrog(torch.full((d,), float(log_eps0), dtype=dtype), L, sc, target))
    Ls.append(Shift(torch.zeros(2 * d, dtype=dtype)))
    Ls.append(Scale(torch.ones(2 * d, dtype=dtype)))
    return Flow(2 * d, Ls, dtype=dtype)


class DiagNormal(Target):
    """MvNormal(mu, Diagonal(sigma^2)) target used by the reference tests (test/objectives.jl:3-6)."""
    kind = "diag_normal"

    def __init__(self, mu, sigma):
        self.mu = torch.as_tensor(mu, dtype=torch.float64)
        self.sigma = torch.as_tensor(sigma, dtype=torch.float64)
        self.dim = self.mu.numel()
        self.params = tuple(self.mu.tolist()) + tuple(self.sigma.tolist())

    def logp(self, y):
        mu, sg = self.mu.to(y.dtype), self.sigma.to(y.dtype)
        z = (y - mu) / sg
        return -0.5 * self.dim * LOG2PI - torch.log(sg).sum() - 0.5 * (z * z).sum(dim=1)


class LogReg(Target):
    """Synthetic Bayesian logistic-regression posterior (BASELINE config 5 names it; the reference itself has no such
    target -- SURVEY section 8f rank 1):  logp(beta) = sum_i [y_i u_i - softplus(u_i)] + log N(beta; 0, sigma0^2 I),
    u = X beta, X [n, dim], y in {0, 1}^n."""
    kind = "logreg"

    def __init__(self, X, y, sigma0=1.0):
        self.X = torch.as_tensor(X, dtype=torch.float64)
        self.y = torch.as_tensor(y, dtype=torch.float64)
        self.sigma0 = float(sigma0)
        self.dim = self.X.shape[1]

    def logp(self, b):
        X, y = self.X.to(b.dtype), self.y.to(b.dtype)
        u = b @ X.T                                             # [N, n]
        ll = (y * u - _softplus(u)).sum(dim=1)
        return ll - 0.5 * (b * b).sum(dim=1) / self.sigma0 ** 2 - 0.5 * self.dim * (LOG2PI + 2 * math.log(self.sigma0))

    def score(self, b):
        X, y = self.X.to(b.dtype), self.y.to(b.dtype)
        u = b @ X.T
        return (y - torch.sigmoid(u)) @ X - b / self.sigma0 ** 2


def synthetic_logreg(dim: int, n: int, seed: int = 7, sigma0: float = 1.0) -> "LogReg":
    """Deterministic synthetic data set: X ~ N(0, 1/dim), beta* ~ N(0, 1), y ~ Bernoulli(sigmoid(X beta*))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.standard_normal((n, dim)) / math.sqrt(dim)
    bstar = rng.standard_normal(dim)
    y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-X @ bstar))).astype(np.float64)
    return LogReg(X, y, sigma0)


# ----------------------------------------------------------------------------------------------
# Objectives
# ----------------------------------------------------------------------------------------------
def batched_elbos(flow: Flow, target: Target, xs: torch.Tensor) -> torch.Tensor:
    """_batched_elbos (src/objectives/elbo.jl:65-70)."""
    ys, ld = flow.forward(xs)
    return target.logp(ys) - flow.base_logpdf(xs) + ld


def elbo_batch(flow: Flow, target: Target, xs: torch.Tensor) -> torch.Tensor:
    """elbo_batch(flow, logp, xs) (src/objectives/elbo.jl:89-92)."""
    return batched_elbos(flow, target, xs).mean()


def elbo(flow: Flow, target: Target, xs: torch.Tensor) -> torch.Tensor:
    """elbo(flow, logp, xs) (src/objectives/elbo.jl:31-34): per-column map then mean."""
    vals = [batched_elbos(flow, target, xs[j:j + 1])[0] for j in range(xs.shape[0])]
    return torch.stack(vals).mean()


def loglikelihood(flow: Flow, xs: torch.Tensor) -> torch.Tensor:
    """loglikelihood(rng, flow, xs) (src/objectives/loglikelihood.jl:26-33)."""
    return flow.logpdf(xs).mean()


def elbo_value_and_grad(flow: Flow, target: Target, theta, xs) -> Tuple[float, np.ndarray]:
    """value and d/dtheta of `elbo_batch(re(theta), logp, xs)` -- what `_value_and_gradient`
    (src/optimize.jl:12-14) returns up to the sign flip of src/NormalizingFlows.jl:69."""
    th = flow.set_theta(torch.as_tensor(theta), requires_grad=True)
    val = elbo_batch(flow, target, xs.to(flow.dtype))
    (g,) = torch.autograd.grad(val, th)
    flow.set_theta(th.detach())
    return float(val.detach()), g.detach().numpy().copy()


def loglik_value_and_grad(flow: Flow, theta, xs) -> Tuple[float, np.ndarray]:
    th = flow.set_theta(torch.as_tensor(theta), requires_grad=True)
    val = loglikelihood(flow, xs.to(flow.dtype))
    (g,) = torch.autograd.grad(val, th)
    flow.set_theta(th.detach())
    return float(val.detach()), g.detach().numpy().copy()


# ----------------------------------------------------------------------------------------------
# Optimisers.Adam + the optimize loop (src/optimize.jl:57-108), for the convergence test
# ----------------------------------------------------------------------------------------------
class Adam:
    """Optimisers.Adam(eta, (0.9, 0.999), 1e-8) (App. A.7)."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), eps=1e-8):
        self.eta, self.b1, self.b2, self.eps = eta, beta[0], beta[1], eps
        self.m = None
        self.v = None
        self.t = 0

    def update(self, theta: np.ndarray, g: np.ndarray) -> np.ndarray:
        if self.m is None:
            self.m = np.zeros_like(theta); self.v = np.zeros_like(theta)
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        mh = self.m / (1 - self.b1 ** self.t)
        vh = self.v / (1 - self.b2 ** self.t)
        return theta - self.eta * mh / (np.sqrt(vh) + self.eps)


def synthetic_z0(n: int, dim: int, seed: int = 2024) -> np.ndarray:
    """Z0 of SURVEY section 8d: Generator(PCG64(seed)).standard_normal((N,d), float32)."""
    return np.random.Generator(np.random.PCG64(seed)).standard_normal((n, dim), dtype=np.float32)
