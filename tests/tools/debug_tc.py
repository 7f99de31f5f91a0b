"""Debug aid (not a pytest): per-parameter-block error of the tcgen05 path vs the fp64 oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import nfload, nf_oracle as O
from helpers import *

nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))
mode = {"simt": nf.NF_MMA_SIMT, "x3": nf.NF_MMA_F16X3, "x1": nf.NF_MMA_F16X1}[sys.argv[1] if len(sys.argv) > 1 else "x3"]

def run(kind, dim, tn, N, kw):
    of = oracle_flow(kind, dim, np.float64, **kw)
    th32 = of.theta().numpy().astype(np.float32)
    of.set_theta(torch.from_numpy(th32).double())
    ot = oracle_target(tn, dim)
    xs = z0(N, dim, np.float32)
    v64, g64 = O.elbo_value_and_grad(of, ot, of.theta(), torch.from_numpy(xs).double())
    gf = gpu_flow(nf, of, np.float32).set_mma_mode(mode)
    gt = gpu_target(nf, ot)
    # forward only
    y, ld = gf.with_logabsdet_jacobian(xs)
    yr, ldr = of.forward(torch.from_numpy(xs).double())
    print(f"[{kind} d={dim} N={N}] fwd y err {rel_err(y, yr.detach().numpy()):.3e}  ld err {rel_err(ld, ldr.detach().numpy()):.3e}")
    v, g = nf.api._elbo_impl(gf, gt, xs, want_grad=True)
    print(f"   elbo {v:.8g} vs {v64:.8g} rel {abs(v-v64)/abs(v64):.3e};  grad rel {rel_err(g, g64):.3e}")
    # per block
    off = 0
    for li, l in enumerate(of.layers):
        mlps = [("s", l.s), ("t", l.t)] if hasattr(l, "s") else [("nn", l.nn)]
        for name, m in mlps:
            for i, (W, b) in enumerate(zip(m.Wts, m.bs)):
                nW, nb = W.numel(), b.numel()
                eW = rel_err(g[off:off+nW], g64[off:off+nW]); off += nW
                eb = rel_err(g[off:off+nb], g64[off:off+nb]); off += nb
                flag = " <<<<" if max(eW, eb) > 1e-3 else ""
                print(f"   L{li}.{name}.dense{i}: W {eW:.2e} b {eb:.2e}{flag}")

run("realnvp", 64, "funnel", 300, dict(hdims=[256, 256], nlayers=1))
run("realnvp", 5, "diag", 64, dict(hdims=[32, 32], nlayers=1))
run("nsf", 16, "cross", 1000, dict(hdims=[32, 32], K=10, B=5.0, nlayers=1))
run("realnvp", 64, "funnel", 5000, dict(hdims=[256, 256], nlayers=4))
