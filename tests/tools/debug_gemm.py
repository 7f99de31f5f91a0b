"""Debug aid: accuracy of the tcgen05 GEMM vs fp64, incl. K-slab accumulation outside the tensor core."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np, ctypes as C
import nfload
nf = nfload.load(); K_ = nf._capi
K_.check(K_.lib().nf_init(0))

def gemm(X, Wt, b, terms=3):
    n, K = X.shape; N = Wt.shape[1]
    Y = np.empty((n, N), np.float32)
    K_.check(K_.lib().nf_tc_gemm_test(n, K, N, K_.ptr(X), K_.ptr(Wt), K_.ptr(b), terms, K_.ptr(Y)))
    return Y

rng = np.random.default_rng(0)
for (n, K, N) in [(1000, 256, 256), (1000, 256, 32), (1000, 32, 256), (333, 64, 232)]:
    X = rng.standard_normal((n, K)).astype(np.float32)
    X = np.where(X > 0, X, 0.01 * X).astype(np.float32)            # leakyrelu-like activations (positive bias)
    Wt = rng.uniform(-0.1, 0.1, (K, N)).astype(np.float32)
    b = np.zeros(N, np.float32)
    ref = X.astype(np.float64) @ Wt.astype(np.float64)
    f32 = X @ Wt
    for terms in (3, 1):
        Y = gemm(X, Wt, b, terms)
        e = (Y - ref)
        print(f"n{n} K{K} N{N} terms{terms}: rel {np.linalg.norm(e)/np.linalg.norm(ref):.3e}  signed bias {np.mean(e*np.sign(ref))/np.mean(np.abs(ref)):+.3e}")
    e = f32 - ref
    print(f"   numpy fp32      : rel {np.linalg.norm(e)/np.linalg.norm(ref):.3e}  signed bias {np.mean(e*np.sign(ref))/np.mean(np.abs(ref)):+.3e}")
    if K >= 128:
        for slab in (64, 128):
            Y = np.zeros((n, N), np.float32)
            for k0 in range(0, K, slab):
                Y += gemm(np.ascontiguousarray(X[:, k0:k0+slab]), np.ascontiguousarray(Wt[k0:k0+slab]), b, 3)
            e = Y - ref
            print(f"   slab {slab:3d} x3 sum : rel {np.linalg.norm(e)/np.linalg.norm(ref):.3e}  signed bias {np.mean(e*np.sign(ref))/np.mean(np.abs(ref)):+.3e}")
# exactly representable inputs: products exact, any error is accumulation
X = (rng.integers(-64, 64, (1000, 256)) / 16.0).astype(np.float32)
Wt = (rng.integers(-64, 64, (256, 256)) / 512.0).astype(np.float32)
ref = X.astype(np.float64) @ Wt.astype(np.float64)
Y = gemm(X, Wt, np.zeros(256, np.float32), 1)
e = Y - ref
print(f"exact-input x1: rel {np.linalg.norm(e)/np.linalg.norm(ref):.3e} signed bias {np.mean(e*np.sign(ref))/np.mean(np.abs(ref)):+.3e}; frac toward zero {np.mean((e*np.sign(ref))<0):.3f} exact {np.mean(e==0):.3f}")
