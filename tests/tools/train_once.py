"""Not a pytest: one persistent-kernel training call of BASELINE config 1 (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np
import nfload
nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))
nf.seed(1)
flow = nf.planarflow(nf.MvNormal(np.zeros(2)), 20, np.float32)
rng = np.random.Generator(np.random.PCG64(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
_, stats, _ = nf.train_flow(rng, nf.elbo, flow, nf.Banana(2, 1.0, 10.0), 10, max_iters=n, optimiser=nf.Adam(1e-3),
                            ADbackend=nf.AutoNFCUDA(on_device=True, chunk=n), show_progress=False)
print("final loss", stats[-1]["loss"])
