"""Run-to-run determinism of the spline bins (and of the ELBO value) with other work interleaved in the same process:
a RealNVP d=64 step (fused coupling kernels, large workspace) runs between repetitions so that stale workspace contents and
different launch timings get a chance to matter.   python tests/tools/bins_repeat.py [reps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import nfload
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, z0
from test_gpu_parity import bin_mismatch_ulps

nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))
d = np.load(os.path.join(ROOT, "tests/golden/c4_nsf_d16_cross.npz"), allow_pickle=False)
meta = json.loads(str(d["meta"]))
of = oracle_flow(meta["kind"], meta["dim"], np.float64, **meta["kw"])
theta32 = d["theta"].astype(np.float32) if "theta" in d.files else of.theta().numpy().astype(np.float32)
of32 = oracle_flow(meta["kind"], meta["dim"], np.float32, **meta["kw"])
of32.set_theta(torch.from_numpy(theta32))
xs = d["z0"].astype(np.float32)
of32.forward(torch.from_numpy(xs))
big = oracle_flow("realnvp", 64, np.float32, hdims=[256, 256], nlayers=2)
gbig = gpu_flow(nf, big, np.float32)
tbig = gpu_target(nf, oracle_target("funnel", 64))
gt = gpu_target(nf, oracle_target(meta["target"], meta["dim"]))
prev, prev_v, bad = None, None, 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for rep in range(reps):
    nf.api._elbo_impl(gbig, tbig, 3000 + 517 * (rep % 5), want_grad=True, seed=rep)
    gf = gpu_flow(nf, of32, np.float32)
    gf.theta = theta32
    got = nf.spline_bins(gf, xs)
    v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    ulps = bin_mismatch_ulps(of32, got)
    same = prev is None or (all(np.array_equal(a, b) for a, b in zip(prev, got)) and v == prev_v)
    if not same or any(u > 32 for u in ulps):
        bad += 1
        print(rep, "mismatches", len(ulps), "ulps", [round(u, 1) for u in ulps][:10], "elbo", v, "identical:", same)
    prev, prev_v = got, v
print("repetitions", reps, "deviating", bad)
