"""Parity table of the BASELINE configurations (small N): CUDA vs the float64 oracle, next to the float32 oracle's own
distance from it, and the spline-bin mismatches with their distance to the separating knot in float32 ulps.

    python tests/tools/parity_table.py [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import nfload
import nf_oracle as O
from helpers import gpu_flow, gpu_target, oracle_flow, oracle_target, rel_err, z0
from test_gpu_parity import CASES

nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))


def bin_mismatches(of, got, dtype):
    """[(layer, ulps)] for every (sample, coordinate) whose bin differs from the oracle's; ulps = distance of the searched
    value from the oracle knot that separates the two bins, in units of the float spacing at that knot."""
    out = []
    for li, (l, g) in enumerate(zip(reversed(of.layers), got)):
        ref = l.last_bins.numpy()
        bad = np.argwhere(g != ref)
        kn, v = l.last_knots.numpy(), l.last_v.numpy()
        for (n, c) in bad:
            lo = min(int(g[n, c]), int(ref[n, c]))
            if abs(int(g[n, c]) - int(ref[n, c])) != 1 or lo < 0 or lo >= kn.shape[-1]:
                out.append((li, float("inf")))
                continue
            knot = kn[n, c, lo]
            out.append((li, float(abs(np.float64(v[n, c]) - np.float64(knot)) / np.spacing(np.float32(abs(knot))))))
    return out


def main(out_path):
    rows = []
    for (kind, dim, tname, N, kw) in CASES:
        of32 = oracle_flow(kind, dim, np.float32, **kw)
        of64 = oracle_flow(kind, dim, np.float64, **kw)
        of64.set_theta(of32.theta().double())
        ot = oracle_target(tname, dim)
        xs = z0(N, dim, np.float32)
        v64, g64 = O.elbo_value_and_grad(of64, ot, of64.theta(), torch.from_numpy(xs).double())
        v32, g32 = O.elbo_value_and_grad(of32, ot, of32.theta(), torch.from_numpy(xs))
        gf = gpu_flow(nf, of32, np.float32)
        v, g = nf.api._elbo_impl(gf, gpu_target(nf, ot), xs, want_grad=True)
        row = dict(config="%s-d%d-%s-N%d" % (kind, dim, tname, N),
                   elbo_rel_gpu_vs_f64=abs(v - v64) / max(abs(v64), 1.0), grad_rel_gpu_vs_f64=rel_err(g, g64),
                   elbo_rel_f32oracle_vs_f64=abs(v32 - v64) / max(abs(v64), 1.0), grad_rel_f32oracle_vs_f64=rel_err(g32, g64),
                   elbo_rel_gpu_vs_f32oracle=abs(v - v32) / max(abs(v32), 1.0), grad_rel_gpu_vs_f32oracle=rel_err(g, g32))
        if kind == "nsf":
            got = nf.spline_bins(gf, xs)
            of64.forward(torch.from_numpy(xs).double())
            mm = bin_mismatches(of64, got, np.float32)
            row.update(bins_total=int(sum(a.size for a in got)), bins_mismatch=len(mm),
                       bins_mismatch_max_ulps=max([u for _, u in mm], default=0.0),
                       bins_mismatch_ulps=sorted(u for _, u in mm))
        row = {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in row.items()}
        rows.append(row)
        print(json.dumps(row))
    with open(out_path, "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_r2.json"))
