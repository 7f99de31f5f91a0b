import os, sys
import numpy as np
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import nfload
from helpers import gpu_flow, oracle_flow, z0
nf = nfload.load(); lib = nf._capi.lib(); nf._capi.check(lib.nf_init(0))
var = int(sys.argv[1]); L = int(sys.argv[2]); N = int(sys.argv[3])
nf._capi.check(lib.nf_set_option(b"fused_variant", var))
of32 = oracle_flow("realnvp", 64, np.float32, hdims=[256, 256], nlayers=L)
gf = gpu_flow(nf, of32, np.float32)
xs = z0(N, 64, np.float32, seed=11)
y1, ld1 = gf.with_logabsdet_jacobian(xs)
for it in range(4):
    y2, ld2 = gf.with_logabsdet_jacobian(xs)
    bad = np.argwhere(y1 != y2)
    print("variant", var, "L", L, "N", N, "run", it, "mismatching elements", len(bad), "rows", len(set(bad[:,0].tolist())), "max abs diff", float(np.abs(y1-y2).max()), "ld maxdiff", float(np.abs(ld1-ld2).max()))
    if len(bad):
        rows = sorted(set(bad[:,0].tolist()))
        print("   first rows", rows[:12], " tiles", sorted(set(r//128 for r in rows))[:12], "cols", sorted(set(bad[:,1].tolist()))[:16])
