"""Not a pytest: batched `rand(flow, n)` / `logpdf(flow, ys)` -- one batched pass through the flow
instead of the per-column loop of reference ext/NormalizingFlowsCUDAExt.jl:65-74.  Host buffers in and out (D2H of the samples
is inside the timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np
import nfload
nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))
nf.seed(123)
for name, flow in (("RealNVP d=64, 8 couplings, 2x256", nf.realnvp(nf.MvNormal(np.zeros(64)), [256, 256], 4, np.float32)),
                   ("NSF d=16, 8 couplings, K=10", nf.nsf(nf.MvNormal(np.zeros(16)), [32, 32], 10, 5.0, 4, np.float32)),
                   ("planar x20 d=2", nf.planarflow(nf.MvNormal(np.zeros(2)), 20, np.float32))):
    import torch
    K = nf._capi
    n = 1 << 20
    ys_t = torch.empty((n, flow.dim), dtype=torch.float32, pin_memory=True)      # page-locked, reused (a fresh pageable numpy array
    lp_t = torch.empty(n, dtype=torch.float32, pin_memory=True)                  #  per call costs more in page faults than the flow)
    ys, lp = ys_t.numpy(), lp_t.numpy()
    th = np.ascontiguousarray(flow.theta)
    K.check(K.lib().nf_sample(flow.handle(), K.ptr(th), n, 1, K.ptr(ys)))
    t0 = time.perf_counter()
    for i in range(5):
        K.check(K.lib().nf_sample(flow.handle(), K.ptr(th), n, 2 + i, K.ptr(ys)))
    t_rand = (time.perf_counter() - t0) / 5
    K.check(K.lib().nf_logpdf(flow.handle(), K.ptr(th), n, K.ptr(ys), K.ptr(lp)))
    t0 = time.perf_counter()
    for i in range(5):
        K.check(K.lib().nf_logpdf(flow.handle(), K.ptr(th), n, K.ptr(ys), K.ptr(lp)))
    t_lp = (time.perf_counter() - t0) / 5
    print("%-36s rand(flow, 2^20): %7.2f ms (%.1f M samples/s)   logpdf(flow, ys): %7.2f ms (%.1f M samples/s)   mean logpdf %.4f"
          % (name, 1e3 * t_rand, n / t_rand / 1e6, 1e3 * t_lp, n / t_lp / 1e6, float(np.mean(lp))))
