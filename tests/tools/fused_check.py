"""Diagnostics for the fused AffineCoupling kernels: fused vs layer-by-layer vs the fp64 oracle, then timing.

    python tests/tools/fused_check.py [--big]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import nfload
import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0

nf = nfload.load()
lib = nf._capi.lib()
nf._capi.check(lib.nf_init(0))


def set_fused(on):
    nf._capi.check(lib.nf_set_option(b"fused_coupling", int(on)))


if "--narrow" in sys.argv:      # two-team streaming kernel (the default)
    nf._capi.check(lib.nf_set_option(b"fused_variant", 0))
if "--wide" in sys.argv:        # 128-column-MMA kernel
    nf._capi.check(lib.nf_set_option(b"fused_variant", 1))


def case(dim, hd, N, nlayers=1, tname="funnel"):
    of32 = oracle_flow("realnvp", dim, np.float32, hdims=hd, nlayers=nlayers)
    of64 = oracle_flow("realnvp", dim, np.float64, hdims=hd, nlayers=nlayers)
    of64.set_theta(of32.theta().double())
    ot = oracle_target(tname, dim)
    xs = z0(N, dim, np.float32)
    v64, g64 = O.elbo_value_and_grad(of64, ot, of64.theta(), torch.from_numpy(xs).double())
    y64, ld64 = of64.forward(torch.from_numpy(xs).double())
    y64 = y64.detach().numpy(); ld64 = ld64.detach().numpy()
    gt = gpu_target(nf, ot)
    out = {}
    for fused in (0, 1):
        set_fused(fused)
        gf = gpu_flow(nf, of32, np.float32)
        try:
            y, ld = gf.with_logabsdet_jacobian(xs)
            v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
        except Exception as e:   # noqa
            print("  fused=%d FAILED: %s" % (fused, e))
            continue
        out[fused] = (y, ld, v, g)
        print("  d=%d hd=%s N=%d L=%d fused=%d: y err %.2e  ld err %.2e  elbo rel %.2e  grad rel %.2e"
              % (dim, hd, N, nlayers, fused, rel_err(y, y64), rel_err(ld, ld64), abs(v - v64) / max(abs(v64), 1.0), rel_err(g, g64)))
    if 0 in out and 1 in out:
        print("     fused vs layered: y %.2e ld %.2e grad %.2e" % (rel_err(out[1][0], out[0][0]), rel_err(out[1][1], out[0][1]), rel_err(out[1][3], out[0][3])))


def timing(N):
    of32 = oracle_flow("realnvp", 64, np.float32, hdims=[256, 256], nlayers=4)
    ot = oracle_target("funnel", 64)
    gt = gpu_target(nf, ot)
    for fused in (0, 1):
        set_fused(fused)
        gf = gpu_flow(nf, of32, np.float32)
        for it in range(4):
            t0 = time.time()
            v, g = nf.api._elbo_impl(gf, gt, N, want_grad=True, seed=5)
            dt = time.time() - t0
        print("  C3 N=%d fused=%d: device %.2f ms (wall %.1f ms) elbo %.6g" % (N, fused, lib.nf_last_device_ms(gf.handle()), dt * 1e3, v))
        import ctypes as C
        lib.nf_profile_enable(gf.handle(), 1)
        v, g = nf.api._elbo_impl(gf, gt, N, want_grad=True, seed=5)
        buf = C.create_string_buffer(4096)
        lib.nf_profile_keys(gf.handle(), buf, 4096)
        for key in buf.value.decode().split(","):
            if not key:
                continue
            cnt, ms = C.c_int64(), C.c_double()
            lib.nf_profile_collect(gf.handle(), key.encode(), C.byref(cnt), C.byref(ms))
            print("      %-28s %3d launches %8.3f ms total %8.4f ms each" % (key, cnt.value, ms.value, ms.value / max(cnt.value, 1)))
        lib.nf_profile_enable(gf.handle(), 0)


if __name__ == "__main__":
    if "--n17" in sys.argv:
        timing(1 << 17)
        sys.exit(0)
    if "--time" in sys.argv:
        timing(1 << 17)
        timing(1 << 20)
        sys.exit(0)
    case(8, [32, 32], 100)
    case(16, [64, 64], 300)
    case(32, [128, 128], 257)
    case(64, [256, 256], 300)
    case(64, [192, 192], 130)
    case(64, [256, 256], 1000, nlayers=4)
    timing(1 << 17)
    if "--big" in sys.argv:
        timing(1 << 20)
