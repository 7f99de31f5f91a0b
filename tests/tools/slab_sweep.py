"""Not a pytest: ELBO / gradient error of the C3-shaped flow against the Float64 oracle for the current NFCUDA_SLAB_FWD /
NFCUDA_RZ_C* environment (accuracy side of the forward slab-drain trade-off), plus the step time at N = 2^20."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import nfload
nf = nfload.load()
import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0
nf._capi.check(nf._capi.lib().nf_init(0))
errs = []
for N, seed in ((1000, 2024), (4096, 7), (4096, 8)):
    of = oracle_flow("realnvp", 64, np.float32, hdims=[256, 256], nlayers=4)
    of64 = oracle_flow("realnvp", 64, np.float64, hdims=[256, 256], nlayers=4)
    of64.set_theta(of.theta().double())
    ot = oracle_target("funnel", 64)
    xs = z0(N, 64, np.float32, seed=seed)
    v64, g64 = O.elbo_value_and_grad(of64, ot, of64.theta(), torch.from_numpy(xs).double())
    v, g = nf.api._elbo_impl(gpu_flow(nf, of, np.float32), gpu_target(nf, ot), xs, want_grad=True)
    errs.append((abs(v - v64) / max(abs(v64), 1.0), rel_err(g, g64)))
print("env", {k: v for k, v in os.environ.items() if k.startswith("NFCUDA_")}, "elbo/grad rel err:", " ".join("%.2e/%.2e" % e for e in errs))
