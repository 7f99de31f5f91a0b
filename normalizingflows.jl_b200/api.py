"""Host-side mirror of the NormalizingFlows.jl interface for the training hot path, over libnfcuda.

Same names, argument meaning and error behaviour as the reference (Julia is not available in this
image, so the host side above the C ABI is Python; the Julia shim that a maintainer would ship is
`julia/NormalizingFlowsNFCUDAExt.jl`):

    train_flow(vo, flow, args...; max_iters, optimiser, ADbackend, ...)   reference src/NormalizingFlows.jl:51-86
    optimize(ad, loss, theta0, re, args...; ...)                          reference src/optimize.jl:57-108
    elbo / elbo_batch / loglikelihood                                     reference src/objectives/*.jl
    planarflow / radialflow / realnvp / nsf / create_flow                 reference src/flows/*.jl
    AffineCoupling / NeuralSplineCoupling / RealNVP_layer / NSF_layer     reference src/flows/realnvp.jl, neuralspline.jl
    destructure(flow) -> (theta, re)                                      Optimisers.destructure (App. A.7)

Differences forced by the host language: batches are numpy `[N, d]` row-major arrays (the memory of a
Julia `d x N` matrix), mask indices are 0-based, and `logp` must be one of the device targets below
(a Python/Julia closure cannot cross the C ABI; `forward_stash`/`backward` cover user densities).
Every numeric result comes from the CUDA library; nothing here computes flows on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _capi as K

Float32, Float64 = np.float32, np.float64


def _dt(paramtype):
    t = np.dtype(paramtype)
    if t == np.float32:
        return K.NF_F32
    if t == np.float64:
        return K.NF_F64
    raise TypeError("paramtype must be Float32 or Float64, got %r" % (paramtype,))


# ------------------------------------------------------------------------------------------------
# base distribution and targets
# ------------------------------------------------------------------------------------------------
class MvNormal:
    """MvNormal(mu, sigma): diagonal Gaussian with STANDARD DEVIATIONS sigma (Distributions' deprecated
    vector form used at reference example/demo_planar_flow.jl:24).  `MvNormal(zeros(d), I)`
    is `MvNormal(np.zeros(d))`."""

    def __init__(self, mu, sigma=None):
        self.mu = np.asarray(mu, dtype=np.float64).reshape(-1)
        self.chol = None
        sg = None if sigma is None else np.asarray(sigma, dtype=np.float64)
        if sg is not None and sg.ndim == 2:
            # MvNormal(mu, Sigma) with a full covariance matrix (reference test/ext/CUDA/cuda.jl:33-37): keep its Cholesky factor
            if sg.shape != (self.mu.size, self.mu.size) or not np.allclose(sg, sg.T):
                raise ValueError("Sigma must be a symmetric dim x dim matrix")
            self.chol = np.ascontiguousarray(np.linalg.cholesky(sg))
            self.sigma = np.sqrt(np.diag(sg))
            return
        self.sigma = np.ones_like(self.mu) if sg is None else sg.reshape(-1)
        if self.sigma.shape != self.mu.shape or np.any(self.sigma <= 0):
            raise ValueError("sigma must be positive with the shape of mu")

    def __len__(self):
        return self.mu.size


class _Target:
    kind = 0
    dim = 0

    def _params(self) -> np.ndarray:
        raise NotImplementedError

    def handle(self):
        if getattr(self, "_h", None) is None:
            p = np.ascontiguousarray(self._params(), dtype=np.float64)
            h = C.c_void_p()
            K.check(K.lib().nf_target_create(C.byref(h), self.kind, self.dim, p.ctypes.data_as(C.POINTER(C.c_double)), p.size))
            self._h = h
        return self._h

    def logp(self, xs, paramtype=np.float64, with_score=False):
        """log-density (and score) of the device target at the rows of xs -- `logpdf(target, x)` of reference example/targets/*.jl."""
        xs = np.ascontiguousarray(np.atleast_2d(xs), dtype=paramtype)
        lp = np.empty(xs.shape[0], dtype=paramtype)
        sc = np.empty_like(xs) if with_score else None
        K.check(K.lib().nf_target_logp(self.handle(), K.NF_F64 if paramtype == np.float64 else K.NF_F32, K.ptr(xs), xs.shape[0],
                                       K.ptr(lp), K.ptr(sc)))
        return (lp, sc) if with_score else lp

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and K is not None and getattr(K, "_lib", None) is not None:   # K is None during interpreter shutdown
            K._lib.nf_target_destroy(h)
            self._h = None


class Banana(_Target):
    """Banana(dim, b, var) -- reference example/targets/banana.jl."""
    kind = K.NF_TARGET_BANANA

    def __init__(self, dim, b, var):
        if dim < 2:
            raise ValueError("dim must be >= 2")
        if var <= 0:
            raise ValueError("var must be > 0")
        self.dim, self.b, self.var = int(dim), float(b), float(var)

    def _params(self):
        return np.array([self.b, self.var])


class Funnel(_Target):
    """Funnel(dim, mu=0, sigma=9) -- reference example/targets/neal_funnel.jl."""
    kind = K.NF_TARGET_FUNNEL

    def __init__(self, dim, mu=0.0, sigma=9.0):
        if dim < 2:
            raise ValueError("dim must be >= 2")
        if sigma <= 0:
            raise ValueError("σ must be > 0")
        self.dim, self.mu, self.sigma = int(dim), float(mu), float(sigma)

    def _params(self):
        return np.array([self.mu, self.sigma])


class WarpedGauss(_Target):
    """WarpedGauss(sigma1=1, sigma2=0.12) -- reference example/targets/warped_gaussian.jl."""
    kind = K.NF_TARGET_WARPED_GAUSS
    dim = 2

    def __init__(self, s1=1.0, s2=0.12):
        if s1 <= 0 or s2 <= 0:
            raise ValueError("σ must be > 0")
        self.s1, self.s2 = float(s1), float(s2)

    def _params(self):
        return np.array([self.s1, self.s2])


class Cross(_Target):
    """Cross(mu=2, sigma=0.15) -- reference example/targets/cross.jl; `dim=2m` gives the product of m
    independent Cross blocks (the synthetic 16-D target of BASELINE config 4)."""
    kind = K.NF_TARGET_CROSS

    def __init__(self, mu=2.0, sigma=0.15, dim=2):
        if dim % 2:
            raise ValueError("dim must be even")
        self.dim, self.mu, self.sigma = int(dim), float(mu), float(sigma)

    def _params(self):
        return np.array([self.mu, self.sigma])


class DiagNormal(_Target):
    """MvNormal(mu, Diagonal(sigma.^2)) as a target (reference test/objectives.jl:3-6, test/ad.jl:41-44)."""
    kind = K.NF_TARGET_DIAG_NORMAL

    def __init__(self, mu, sigma):
        self.mu = np.asarray(mu, dtype=np.float64).reshape(-1)
        self.sigma = np.asarray(sigma, dtype=np.float64).reshape(-1)
        self.dim = self.mu.size

    def _params(self):
        return np.concatenate([self.mu, self.sigma])


class LogReg(_Target):
    """Synthetic Bayesian logistic-regression posterior over `dim` coefficients (BASELINE config 5; not in the reference):
    logp(b) = sum_i [y_i x_i.b - softplus(x_i.b)] + log N(b; 0, sigma0^2 I) for X [n, dim], y in {0,1}^n.  The data set is
    copied to the device once; logp, score and Hessian-vector products are evaluated inside the kernels."""
    kind = K.NF_TARGET_LOGREG

    def __init__(self, X, y, sigma0=1.0):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        if self.X.ndim != 2 or self.X.shape[0] != self.y.size:
            raise ValueError("X must be [n, dim] and y of length n")
        if sigma0 <= 0:
            raise ValueError("sigma0 must be > 0")
        self.dim, self.sigma0 = int(self.X.shape[1]), float(sigma0)

    def _params(self):
        return np.concatenate([[self.sigma0, float(self.y.size)], self.X.reshape(-1), self.y])


class JointTarget(_Target):
    """logp_joint(z) = logp(x) + sum(logpdf(Normal(), rho)) on z = [x, rho]
    (reference example/demo_hamiltonian_flow.jl:117-124)."""

    def __init__(self, inner: _Target):
        self.inner, self.dim = inner, 2 * inner.dim

    def handle(self):
        if getattr(self, "_h", None) is None:
            h = C.c_void_p()
            K.check(K.lib().nf_target_create_joint(C.byref(h), self.inner.handle()))
            self._h = h
        return self._h


# ------------------------------------------------------------------------------------------------
# layers (structure + initial parameters; all arithmetic lives in the CUDA library)
# ------------------------------------------------------------------------------------------------
_RNG = np.random.Generator(np.random.PCG64(123))


def seed(s: int):
    """Seed the host RNG used by layer constructors (stands in for Random.seed!)."""
    global _RNG
    _RNG = np.random.Generator(np.random.PCG64(s))


@dataclass
class _Layer:
    kind: int
    dim: int
    theta: np.ndarray                       # float64 master copy in destructure order
    mask_idx: Optional[List[int]] = None
    hdims: Optional[List[int]] = None
    K: int = 0
    B: float = 0.0
    n_steps: int = 0
    score_target: Optional[object] = None

    def __matmul__(self, other):            # l1 @ l2  ==  l1 ∘ l2
        return Composed(_flatten(self) + _flatten(other))


@dataclass
class Composed:
    """ComposedFunction: `layers` in theta order; applied last-to-first."""
    layers: List[_Layer]

    def __matmul__(self, other):
        return Composed(self.layers + _flatten(other))


def _flatten(x) -> List[_Layer]:
    if isinstance(x, Composed):
        return list(x.layers)
    if isinstance(x, _Layer):
        return [x]
    raise TypeError("not a bijector layer: %r" % (x,))


def PlanarLayer(dim: int) -> _Layer:
    """Bijectors.PlanarLayer(dim): w, u, b ~ randn (App. A.1)."""
    return _Layer(K.NF_PLANAR, dim, _RNG.standard_normal(2 * dim + 1))


def RadialLayer(dim: int) -> _Layer:
    """Bijectors.RadialLayer(dim): alpha_, beta, z0 ~ randn (App. A.2)."""
    return _Layer(K.NF_RADIAL, dim, _RNG.standard_normal(dim + 2))


def Shift(a) -> _Layer:
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return _Layer(K.NF_SHIFT, a.size, a.copy())


def Scale(a) -> _Layer:
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return _Layer(K.NF_SCALE, a.size, a.copy())


def MomentumAffine(shift, scale) -> _Layer:
    """`Stacked((identity, Shift(b) ∘ Scale(a)), [1:h, h+1:2h])`: the momentum normalisation layer of
    reference example/demo_hamiltonian_flow.jl:94-99 on z = [x, rho]; theta = [b; a]."""
    b = np.asarray(shift, dtype=np.float64).reshape(-1)
    a = np.asarray(scale, dtype=np.float64).reshape(-1)
    if a.shape != b.shape:
        raise ValueError("shift and scale must have the same length")
    return _Layer(K.NF_MOMENTUM_AFFINE, 2 * a.size, np.concatenate([b, a]))


def LeapFrog(dim: int, log_eps, L: int, target) -> _Layer:
    """LeapFrog(dim, logϵ, L, ∇logp) -- reference example/demo_hamiltonian_flow.jl:27-47: `dim` position coordinates,
    per-dimension log step sizes (trainable), L leapfrog steps, the score of `target` (a device target) inside."""
    le = np.asarray(log_eps, dtype=np.float64).reshape(-1)
    if le.size == 1:
        le = np.full(dim, float(le[0]))
    if le.size != dim or target.dim != dim:
        raise ValueError("log_eps and target must have dimension %d" % dim)
    return _Layer(K.NF_LEAPFROG, 2 * dim, le.copy(), n_steps=int(L), score_target=target)


def _fnn_theta(n_in: int, hdims: Sequence[int], n_out: int) -> np.ndarray:
    """fnn (reference src/flows/utils.jl:71-100) initial parameters: Dense weight ~ glorot_uniform drawn
    in Float32, bias zeros (App. A.6), flattened as [W1(:); b1; W2(:); b2; ...] with W out x in column major."""
    dims = [n_in] + list(hdims) + [n_out]
    parts = []
    for a, b in zip(dims[:-1], dims[1:]):
        lim = math.sqrt(6.0 / (a + b))
        W = _RNG.uniform(-lim, lim, size=(a, b)).astype(np.float32)   # [in, out] row-major == vec(out x in)
        parts += [W.reshape(-1).astype(np.float64), np.zeros(b)]
    return np.concatenate(parts)


def AffineCoupling(dim: int, hdims: Sequence[int], mask_idx: Sequence[int], paramtype=Float64) -> _Layer:
    """AffineCoupling(dim, hdims, mask_idx, paramtype) -- reference src/flows/realnvp.jl:42-55.
    `mask_idx` (0-based) are the transformed coordinates."""
    _dt(paramtype)
    c = len(mask_idx)
    th = np.concatenate([_fnn_theta(dim - c, hdims, c), _fnn_theta(dim - c, hdims, c)])
    return _Layer(K.NF_AFFINE_COUPLING, dim, th, list(map(int, mask_idx)), list(map(int, hdims)))


def NeuralSplineCoupling(dim: int, hdims: Sequence[int], K_: int, B: float, mask_idx: Sequence[int], paramtype=Float64) -> _Layer:
    """NeuralSplineCoupling(dim, hdims, K, B, mask_idx, paramtype) -- reference src/flows/neuralspline.jl:44-61."""
    _dt(paramtype)
    c = len(mask_idx)
    th = _fnn_theta(dim - c, hdims, (3 * K_ - 1) * c)
    return _Layer(K.NF_SPLINE_COUPLING, dim, th, list(map(int, mask_idx)), list(map(int, hdims)), int(K_), float(B))


def RealNVP_layer(dims: int, hdims: Sequence[int], paramtype=Float64) -> Composed:
    """af(1:2:d) ∘ af(2:2:d) -- reference src/flows/realnvp.jl:132-145."""
    af1 = AffineCoupling(dims, hdims, list(range(0, dims, 2)), paramtype)
    af2 = AffineCoupling(dims, hdims, list(range(1, dims, 2)), paramtype)
    return af1 @ af2


def NSF_layer(dim: int, hdims: Sequence[int], K_: int, B: float, paramtype=Float64) -> Composed:
    """nsc(1:2:d) ∘ nsc(2:2:d) -- reference src/flows/neuralspline.jl:169-184."""
    n1 = NeuralSplineCoupling(dim, hdims, K_, B, list(range(0, dim, 2)), paramtype)
    n2 = NeuralSplineCoupling(dim, hdims, K_, B, list(range(1, dim, 2)), paramtype)
    return n1 @ n2


# ------------------------------------------------------------------------------------------------
# Flow == Bijectors.TransformedDistribution over the CUDA handle
# ------------------------------------------------------------------------------------------------
class Flow:
    """transformed(q0, reduce(∘, Ls)) -- reference src/flows/utils.jl:23-26."""

    def __init__(self, layers: List[_Layer], q0: MvNormal, paramtype=Float64, theta: Optional[np.ndarray] = None):
        if not layers:
            raise ValueError("a flow needs at least one layer")
        self.layers = layers
        self.dist = q0
        self.dim = len(q0)
        for l in layers:
            if l.dim != self.dim:
                raise ValueError("layer dim %d != base dim %d" % (l.dim, self.dim))
        self.paramtype = np.dtype(paramtype).type
        self.theta = (np.concatenate([l.theta for l in layers]) if theta is None else np.asarray(theta)).astype(self.paramtype)
        self._h = None
        self._keep = None
        self._owner = None

    # -- handle ---------------------------------------------------------------------------------
    def handle(self):
        owner = self.__dict__.get("_owner")
        if owner is not None:          # `re(theta)` views share the owner's CUDA handle
            return owner.handle()
        if self._h is None:
            n = len(self.layers)
            descs = (K.LayerDesc * n)()
            keep = []
            for i, l in enumerate(self.layers):
                descs[i].kind = l.kind
                if l.mask_idx is not None:
                    m = (C.c_int * len(l.mask_idx))(*l.mask_idx)
                    h = (C.c_int * len(l.hdims))(*l.hdims)
                    keep += [m, h]
                    descs[i].mask_idx = m; descs[i].n_mask = len(l.mask_idx)
                    descs[i].hdims = h; descs[i].n_hidden = len(l.hdims)
                descs[i].K = l.K
                descs[i].B = l.B
                descs[i].n_steps = l.n_steps
                if l.score_target is not None:
                    descs[i].score_target = l.score_target.handle()
                    keep.append(l.score_target)
            h = C.c_void_p()
            K.check(K.lib().nf_flow_create(C.byref(h), descs, n, self.dim, _dt(self.paramtype)))
            self._h = h
            self._keep = keep
            P = K.lib().nf_flow_num_params(h)
            if P != self.theta.size:
                raise K.NFCudaError("parameter count mismatch: library %d vs host %d" % (P, self.theta.size))
            mu = np.ascontiguousarray(self.dist.mu); sg = np.ascontiguousarray(self.dist.sigma)
            if getattr(self.dist, "chol", None) is not None:
                Lc = np.ascontiguousarray(self.dist.chol, dtype=np.float64)
                K.check(K.lib().nf_flow_set_base_chol(h, mu.ctypes.data_as(C.POINTER(C.c_double)), Lc.ctypes.data_as(C.POINTER(C.c_double))))
            else:
                K.check(K.lib().nf_flow_set_base(h, mu.ctypes.data_as(C.POINTER(C.c_double)), sg.ctypes.data_as(C.POINTER(C.c_double))))
        return self._h

    def __del__(self):
        if getattr(self, "_h", None) is not None and K is not None and getattr(K, "_lib", None) is not None:
            K._lib.nf_flow_destroy(self._h)
            self._h = None

    def set_mma_mode(self, mode: int):
        K.check(K.lib().nf_flow_set_mma_mode(self.handle(), mode))
        return self

    def set_workspace_limit(self, nbytes: int):
        K.check(K.lib().nf_flow_set_workspace_limit(self.handle(), nbytes))
        return self

    @property
    def num_params(self):
        return self.theta.size

    def _x(self, xs):
        xs = np.ascontiguousarray(xs, dtype=self.paramtype)
        if xs.ndim == 1:
            xs = xs.reshape(1, -1)
        if xs.ndim != 2 or xs.shape[1] != self.dim:
            raise ValueError("expected a [N, %d] batch (memory of a Julia %d x N matrix)" % (self.dim, self.dim))
        return xs

    # -- Bijectors / Distributions surface ---------------------------------------------------------
    def with_logabsdet_jacobian(self, xs):
        """with_logabsdet_jacobian(flow.transform, xs) -> (ys, logabsdetjac)."""
        xs = self._x(xs)
        ys = np.empty_like(xs); ld = np.empty(xs.shape[0], dtype=self.paramtype)
        K.check(K.lib().nf_forward(self.handle(), K.ptr(self.theta), xs.shape[0], K.ptr(xs), K.ptr(ys), K.ptr(ld)))
        return ys, ld

    def inverse_with_logabsdet_jacobian(self, ys):
        """with_logabsdet_jacobian(inverse(flow.transform), ys) -> (xs, logabsdetjac)."""
        ys = self._x(ys)
        xs = np.empty_like(ys); ld = np.empty(ys.shape[0], dtype=self.paramtype)
        K.check(K.lib().nf_inverse(self.handle(), K.ptr(self.theta), ys.shape[0], K.ptr(ys), K.ptr(xs), K.ptr(ld)))
        return xs, ld

    def logpdf(self, ys, out=None):
        """logpdf(flow, ys) -- one value per sample.  `out`: optional preallocated (ideally page-locked) result array."""
        ys = self._x(ys)
        if out is None:
            out = np.empty(ys.shape[0], dtype=self.paramtype)
        elif out.shape != (ys.shape[0],) or out.dtype != self.paramtype or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous %s array of shape (%d,)" % (np.dtype(self.paramtype).name, ys.shape[0]))
        K.check(K.lib().nf_logpdf(self.handle(), K.ptr(self.theta), ys.shape[0], K.ptr(ys), K.ptr(out)))
        return out

    def rand(self, n: int, seed: Optional[int] = None, out=None):
        """rand(flow, n): device Philox base draws pushed through the flow in one batched pass.  `out`: optional preallocated
        [n, dim] array (a fresh pageable array per call costs more in page faults than the flow itself at n = 2^20)."""
        if out is None:
            ys = np.empty((n, self.dim), dtype=self.paramtype)
        elif out.shape != (n, self.dim) or out.dtype != self.paramtype or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous %s array of shape (%d, %d)" % (np.dtype(self.paramtype).name, n, self.dim))
        else:
            ys = out
        if seed is None:                      # fresh draws per call, like rand(flow, n) advancing the RNG; pass a seed to reproduce
            seed = int(_RNG.integers(0, 2 ** 63))
        K.check(K.lib().nf_sample(self.handle(), K.ptr(self.theta), n, seed, K.ptr(ys)))
        return ys

    def rand_base(self, n: int, seed: Optional[int] = None):
        """_device_specific_rand(rng, flow.dist, n) -- reference src/NormalizingFlows.jl:109-115."""
        zs = np.empty((n, self.dim), dtype=self.paramtype)
        if seed is None:
            seed = int(_RNG.integers(0, 2 ** 63))
        K.check(K.lib().nf_base_sample(self.handle(), n, seed, K.ptr(zs)))
        return zs


def create_flow(Ls, q0: MvNormal, paramtype=Float64) -> Flow:
    """create_flow(Ls, q0) = transformed(q0, reduce(∘, Ls)) -- reference src/flows/utils.jl:23-26."""
    layers: List[_Layer] = []
    for L in Ls:
        layers += _flatten(L)
    return Flow(layers, q0, paramtype)


def transformed(q0: MvNormal, bijector, paramtype=Float64) -> Flow:
    """Bijectors.transformed(q0, b)."""
    return Flow(_flatten(bijector), q0, paramtype)


def planarflow(q0: MvNormal, nlayers: int, paramtype=Float64) -> Flow:
    """planarflow(q0, nlayers; paramtype) -- reference src/flows/planar_radial.jl:21-29."""
    return create_flow([PlanarLayer(len(q0)) for _ in range(nlayers)], q0, paramtype)


def radialflow(q0: MvNormal, nlayers: int, paramtype=Float64) -> Flow:
    """radialflow(q0, nlayers; paramtype) -- reference src/flows/planar_radial.jl:52-60."""
    return create_flow([RadialLayer(len(q0)) for _ in range(nlayers)], q0, paramtype)


def realnvp(q0: MvNormal, hdims: Sequence[int] = (32, 32), nlayers: int = 10, paramtype=Float64) -> Flow:
    """realnvp(q0, hdims, nlayers; paramtype) -- reference src/flows/realnvp.jl:170-192."""
    return create_flow([RealNVP_layer(len(q0), hdims, paramtype) for _ in range(nlayers)], q0, paramtype)


def nsf(q0: MvNormal, hdims: Sequence[int] = (32, 32), K_: int = 10, B: float = 30.0, nlayers: int = 10, paramtype=Float64) -> Flow:
    """nsf(q0, hdims, K, B, nlayers; paramtype) -- reference src/flows/neuralspline.jl:218-234."""
    return create_flow([NSF_layer(len(q0), hdims, K_, B, paramtype) for _ in range(nlayers)], q0, paramtype)


def hamiltonian_flow(target, nlayers: int = 15, L: int = 3, log_eps0: float = math.log(0.05), paramtype=Float64) -> Flow:
    """The Hamiltonian flow of reference example/demo_hamiltonian_flow.jl:128-147 for a device target over `dims`
    coordinates: q0 = transformed(MvNormal(0, I_2dims), Shift(0) ∘ Scale(1)) and
    Ls = [momentum_normalization_layer ∘ LeapFrog(dims, log ϵ0, L, ∇logp) for _ in 1:nlayers].  Train it against
    `JointTarget(target)`.  theta order (Optimisers.destructure of `transformed(q0, ∘(Ls...))`): per layer
    [b; a; logϵ], then q0's shift, scale -- which are applied first (App. A.5)."""
    h = target.dim
    layers: List[_Layer] = []
    for _ in range(nlayers):
        layers += _flatten(MomentumAffine(np.zeros(h), np.ones(h)) @ LeapFrog(h, log_eps0, L, target))
    layers += _flatten(Shift(np.zeros(2 * h)) @ Scale(np.ones(2 * h)))
    return Flow(layers, MvNormal(np.zeros(2 * h)), paramtype)


def destructure(flow: Flow):
    """Optimisers.destructure(flow) -> (theta_flat, re) -- reference src/NormalizingFlows.jl:67."""
    def re(theta):
        theta = np.asarray(theta)
        if theta.size != flow.theta.size:
            raise ValueError("theta has %d elements, flow has %d parameters" % (theta.size, flow.theta.size))
        g = Flow.__new__(Flow)
        g.__dict__.update(flow.__dict__)
        g.theta = theta.astype(flow.paramtype)
        g._owner = flow.__dict__.get("_owner") or flow   # shares the CUDA handle; keeps the owner alive
        g._h = None
        return g
    return flow.theta.copy(), re



# ------------------------------------------------------------------------------------------------
# objectives (reference src/objectives/elbo.jl, loglikelihood.jl)
# ------------------------------------------------------------------------------------------------
def _elbo_impl(flow: Flow, logp: _Target, xs_or_n, rng=None, want_grad=False, scale=1.0, seed=None):
    if not isinstance(logp, _Target):
        raise TypeError("logp must be a device target (Banana, Funnel, WarpedGauss, Cross, DiagNormal); "
                        "use forward_stash/backward for a user-supplied density")
    val = C.c_double()
    grad = np.empty(flow.theta.size, dtype=flow.paramtype) if want_grad else None
    if isinstance(xs_or_n, (int, np.integer)):
        n = int(xs_or_n)
        sd = seed if seed is not None else (int(rng.integers(0, 2 ** 63)) if rng is not None else int(_RNG.integers(0, 2 ** 63)))
        K.check(K.lib().nf_elbo_value_and_grad(flow.handle(), logp.handle(), K.ptr(flow.theta), n, None, sd, scale,
                                               C.byref(val), K.ptr(grad)))
    else:
        xs = flow._x(xs_or_n)
        K.check(K.lib().nf_elbo_value_and_grad(flow.handle(), logp.handle(), K.ptr(flow.theta), xs.shape[0], K.ptr(xs), 0,
                                               scale, C.byref(val), K.ptr(grad)))
    v = flow.paramtype(val.value)
    return (v, grad) if want_grad else v


def elbo(*args):
    """elbo(flow, logp, xs) | elbo([rng,] flow, logp, n_samples) -- reference src/objectives/elbo.jl:26-46.
    The per-column `map` of the reference and the batched pass are the same arithmetic on the GPU."""
    rng, args = (args[0], args[1:]) if isinstance(args[0], np.random.Generator) else (None, args)
    flow, logp, xs_or_n = args
    return _elbo_impl(flow, logp, xs_or_n, rng)


def elbo_batch(*args):
    """elbo_batch(flow, logp, xs) | elbo_batch([rng,] flow, logp, n_samples) -- reference elbo.jl:89-99."""
    return elbo(*args)


def batched_elbos(flow: Flow, logp: _Target, xs):
    """_batched_elbos(flow, logp, xs) -- reference src/objectives/elbo.jl:65-70."""
    xs = flow._x(xs)
    out = np.empty(xs.shape[0], dtype=flow.paramtype)
    K.check(K.lib().nf_elbo_terms(flow.handle(), logp.handle(), K.ptr(flow.theta), xs.shape[0], K.ptr(xs), K.ptr(out)))
    return out


def loglikelihood(*args):
    """loglikelihood(rng, flow, xs) -- reference src/objectives/loglikelihood.jl:18-33 (rng is ignored)."""
    if isinstance(args[0], np.random.Generator) or args[0] is None:
        args = args[1:]
    flow, xs = args
    xs = flow._x(xs)
    val = C.c_double()
    K.check(K.lib().nf_loglik_value_and_grad(flow.handle(), K.ptr(flow.theta), xs.shape[0], K.ptr(xs), 1.0, C.byref(val), None))
    return flow.paramtype(val.value)


# ------------------------------------------------------------------------------------------------
# AD seam + optimisation loop (reference src/optimize.jl)
# ------------------------------------------------------------------------------------------------
class AutoNFCUDA:
    """The ADTypes-style backend tag a Julia user would pass as `ADbackend`.

    `on_device=True` additionally keeps the optimiser step on the GPU (`nf_train_elbo_adam`): used by `optimize` when the
    objective is elbo/elbo_batch with a sample count, the optimiser is Adam and there is neither a callback nor a
    convergence test that needs theta on the host every iteration."""

    def __init__(self, on_device: bool = False, chunk: int = 100):
        self.on_device, self.chunk = on_device, chunk

    def __repr__(self):
        return "AutoNFCUDA(on_device=%r)" % self.on_device


@dataclass
class _Loss:
    """loss(theta, rng, args...) = -vo(rng, re(theta), args...) -- reference src/NormalizingFlows.jl:69."""
    vo: Callable
    re: Callable

    def __call__(self, theta, rng, *args):
        return -self.vo(rng, self.re(theta), *args)


def _prepare_gradient(loss: _Loss, adbackend, theta, *args):
    """reference src/optimize.jl:8-10: one-off preparation = creating the CUDA flow handle."""
    if not isinstance(adbackend, AutoNFCUDA):
        raise TypeError("only ADbackend=AutoNFCUDA() is available in this build")
    flow = loss.re(theta)
    flow.handle()
    return flow


def _value_and_gradient(loss: _Loss, prep, adbackend, theta, *args):
    """reference src/optimize.jl:12-14: (loss value, gradient) in one library call."""
    flow = loss.re(theta)
    rng, rest = args[0], args[1:]
    if loss.vo in (elbo, elbo_batch):
        logp, n_or_xs = rest
        return _elbo_impl(flow, logp, n_or_xs, rng, want_grad=True, scale=-1.0)
    if loss.vo is loglikelihood:
        (xs,) = rest
        xs = flow._x(xs)
        val = C.c_double()
        grad = np.empty(flow.theta.size, dtype=flow.paramtype)
        K.check(K.lib().nf_loglik_value_and_grad(flow.handle(), K.ptr(flow.theta), xs.shape[0], K.ptr(xs), -1.0,
                                                 C.byref(val), K.ptr(grad)))
        return flow.paramtype(val.value), grad
    raise TypeError("variational objective must be elbo, elbo_batch or loglikelihood")


class Adam:
    """Optimisers.Adam(eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8) -- stays on the host like Optimisers.update!
    (reference src/optimize.jl:99)."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8):
        self.eta, self.beta, self.epsilon = eta, beta, epsilon

    def setup(self, theta):
        return {"m": np.zeros_like(theta), "v": np.zeros_like(theta), "bt": (self.beta[0], self.beta[1])}

    def update(self, st, theta, g):
        b1, b2 = self.beta
        dt = theta.dtype.type
        st["m"] = dt(b1) * st["m"] + dt(1 - b1) * g
        st["v"] = dt(b2) * st["v"] + dt(1 - b2) * g * g
        p1, p2 = st["bt"]
        step = st["m"] / dt(1 - p1) / (np.sqrt(st["v"] / dt(1 - p2)) + dt(self.epsilon)) * dt(self.eta)
        st["bt"] = (p1 * b1, p2 * b2)
        return st, theta - step


ADAM = Adam


def optimize(adbackend, loss, theta0, reconstruct, *args, max_iters=10000, optimiser=None, show_progress=True,
             callback=None, hasconverged=None, prog=None):
    """optimize(ad, loss, theta0, re, args...; ...) -- reference src/optimize.jl:57-108."""
    optimiser = optimiser or Adam()
    if (getattr(adbackend, "on_device", False) and callback is None and hasconverged is None and isinstance(optimiser, Adam)
            and isinstance(loss, _Loss) and loss.vo in (elbo, elbo_batch) and len(args) == 3 and isinstance(args[2], (int, np.integer))):
        return _optimize_on_device(adbackend, loss, theta0, reconstruct, *args, max_iters=max_iters, optimiser=optimiser,
                                   show_progress=show_progress)
    hasconverged = hasconverged or (lambda i, stats, re, theta, st: False)
    opt_stats = []
    theta = np.array(theta0, copy=True)
    prep = _prepare_gradient(loss, adbackend, theta0, *args)
    st = optimiser.setup(theta)
    converged = False
    i = 1
    t0 = time.time()
    while i <= max_iters and not converged:
        ls, g = _value_and_gradient(loss, prep, adbackend, theta, *args)
        stat = {"iteration": i, "loss": ls, "gradient_norm": float(np.linalg.norm(g))}
        if callback is not None:
            new_stat = callback(i, opt_stats, reconstruct, theta)
            if new_stat is not None:
                stat.update(new_stat)
        opt_stats.append(stat)
        st, theta = optimiser.update(st, theta, g)
        i += 1
        converged = hasconverged(i, stat, reconstruct, theta, st)
        if show_progress and (i % 100 == 0 or converged or i > max_iters):
            print("Training %d/%d  loss=%.6g  |g|=%.4g  (%.1f it/s)" % (i - 1, max_iters, ls, stat["gradient_norm"],
                                                                      (i - 1) / max(time.time() - t0, 1e-9)))
    return theta, opt_stats, st


def _optimize_on_device(adbackend, loss, theta0, reconstruct, rng, logp, n, max_iters, optimiser, show_progress):
    """The same loop with the Adam step fused on the device; theta visits the host once per `chunk` iterations."""
    flow = loss.re(theta0)
    dt = flow.paramtype
    theta = np.array(theta0, dtype=dt, copy=True)
    m = np.zeros_like(theta); v = np.zeros_like(theta)
    opt_stats, done = [], 0
    t_start = time.time()
    while done < max_iters:
        k = min(adbackend.chunk, max_iters - done)
        stats = np.empty((k, 2), dtype=np.float64)
        seed = int(rng.integers(0, 2 ** 62))
        K.check(K.lib().nf_train_elbo_adam(flow.handle(), logp.handle(), K.ptr(theta), int(n), seed, k, done, float(optimiser.eta),
                                           float(optimiser.beta[0]), float(optimiser.beta[1]), float(optimiser.epsilon),
                                           K.ptr(m), K.ptr(v), stats.ctypes.data_as(C.POINTER(C.c_double))))
        for i in range(k):
            opt_stats.append({"iteration": done + i + 1, "loss": dt(stats[i, 0]), "gradient_norm": float(stats[i, 1])})
        done += k
        if show_progress:
            print("Training %d/%d  loss=%.6g  |g|=%.4g  (%.1f it/s, on-device Adam)" % (done, max_iters, stats[-1, 0], stats[-1, 1],
                                                                                       done / max(time.time() - t_start, 1e-9)))
    b1, b2 = optimiser.beta
    st = {"m": m, "v": v, "bt": (b1 ** (done + 1), b2 ** (done + 1))}
    return theta, opt_stats, st


def train_flow(*args, max_iters=1000, optimiser=None, ADbackend=None, **kwargs):
    """train_flow([rng,] vo, flow, args...; max_iters, optimiser, ADbackend, kwargs...) -- reference
    src/NormalizingFlows.jl:51-86.  `ADbackend` is required (no default), as in the reference (:61)."""
    if ADbackend is None:
        raise TypeError("train_flow: keyword argument ADbackend not assigned")
    if isinstance(args[0], np.random.Generator):
        rng, args = args[0], args[1:]
    else:
        rng = np.random.default_rng()
    vo, flow, rest = args[0], args[1], args[2:]
    theta_flat, re = destructure(flow)
    loss = _Loss(vo, re)
    theta_trained, opt_stats, st = optimize(ADbackend, loss, theta_flat, re, rng, *rest, max_iters=max_iters,
                                            optimiser=optimiser or Adam(), **kwargs)
    return re(theta_trained), opt_stats, st


# ------------------------------------------------------------------------------------------------
# two-phase API for user-supplied log-densities (SURVEY section 7 'Arbitrary Julia logp')
# ------------------------------------------------------------------------------------------------
def forward_stash(flow: Flow, xs):
    xs = flow._x(xs)
    ys = np.empty_like(xs); ld = np.empty(xs.shape[0], dtype=flow.paramtype)
    K.check(K.lib().nf_forward_stash(flow.handle(), K.ptr(flow.theta), xs.shape[0], K.ptr(xs), K.ptr(ys), K.ptr(ld)))
    return ys, ld


def backward(flow: Flow, gy, gld=None):
    gy = flow._x(gy)
    gld = None if gld is None else np.ascontiguousarray(gld, dtype=flow.paramtype)
    grad = np.empty(flow.theta.size, dtype=flow.paramtype)
    K.check(K.lib().nf_backward(flow.handle(), K.ptr(gy), K.ptr(gld), K.ptr(grad)))
    return grad


def spline_bins(flow: Flow, xs):
    """Bin index (searchsortedfirst - 1) of every transformed coordinate in every spline coupling, in
    application order: list of int32 [N, c] arrays."""
    xs = flow._x(xs)
    sizes = [len(l.mask_idx) for l in reversed(flow.layers) if l.kind == K.NF_SPLINE_COUPLING]
    total = sum(sizes) * xs.shape[0]
    out = np.empty(total, dtype=np.int32)
    K.check(K.lib().nf_spline_bins(flow.handle(), K.ptr(flow.theta), xs.shape[0], K.ptr(xs), out.ctypes.data_as(C.POINTER(C.c_int32))))
    res, off = [], 0
    for c in sizes:
        res.append(out[off:off + xs.shape[0] * c].reshape(xs.shape[0], c))
        off += xs.shape[0] * c
    return res


def rqs_bin_search(knots, v):
    knots = np.ascontiguousarray(knots); v = np.ascontiguousarray(v, dtype=knots.dtype)
    M, K1 = knots.shape
    out = np.empty(M, dtype=np.int32)
    K.check(K.lib().nf_rqs_bin_search(_dt(knots.dtype), K.ptr(knots), K.ptr(v), M, K1 - 1, out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out
