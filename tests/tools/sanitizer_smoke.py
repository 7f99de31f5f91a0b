"""Not a pytest: a tiny pass over every kernel family for `compute-sanitizer --tool memcheck`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import nfload
nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))
rng = np.random.default_rng(0)
for T in (np.float32, np.float64):
    for name, flow, tgt in [
        ("planar", nf.planarflow(nf.MvNormal(np.zeros(2)), 5, T), nf.Banana(2, 1.0, 10.0)),
        ("radial", nf.radialflow(nf.MvNormal(np.zeros(3)), 4, T), nf.DiagNormal(np.zeros(3), np.ones(3))),
        ("realnvp", nf.realnvp(nf.MvNormal(np.zeros(6)), [16, 16], 1, T), nf.Funnel(6)),
        ("realnvp64", nf.realnvp(nf.MvNormal(np.zeros(64)), [256, 256], 1, T), nf.Funnel(64)),
        ("nsf", nf.nsf(nf.MvNormal(np.zeros(4)), [8, 8], 5, 3.0, 1, T), nf.Cross(2.0, 0.15, 4)),
        ("nsf16", nf.nsf(nf.MvNormal(np.zeros(16)), [32, 32], 10, 5.0, 1, T), nf.Cross(2.0, 0.15, 16)),   # K=10 bulk-copy ring, 1 tile + ragged tail
        ("hamiltonian", nf.hamiltonian_flow(nf.Funnel(2, -2.0, 3.0), 3, 2, -3.0, T), nf.JointTarget(nf.Funnel(2, -2.0, 3.0))),
    ]:
        if name == "realnvp64" and T == np.float64:
            continue
        xs = rng.standard_normal((333 if name == 'nsf16' else 200, flow.dim)).astype(T)
        v, g = nf.api._elbo_impl(flow, tgt, xs, want_grad=True)
        y, ld = flow.with_logabsdet_jacobian(xs)
        x2, ld2 = flow.inverse_with_logabsdet_jacobian(y)
        ll = nf.loglikelihood(None, flow, y)
        print(name, T.__name__, float(v), float(np.linalg.norm(g)), float(np.abs(x2 - xs).max()), float(ll))
print("sanitizer smoke done")
