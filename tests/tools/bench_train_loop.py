"""Not a pytest: iterations/s of the training loop for BASELINE config 1 (planar x20, d=2, Banana, batch 10) -- the
latency-bound demo shape -- host loop vs on-device Adam loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np
import nfload
nf = nfload.load()
nf._capi.check(nf._capi.lib().nf_init(0))
for on_dev, persistent in ((False, 1), (True, 0), (True, 1)):
    os.environ["NFCUDA_TRAIN_PERSISTENT"] = str(persistent)
    nf.seed(1)
    flow = nf.planarflow(nf.MvNormal(np.zeros(2)), 20, np.float32)
    target = nf.Banana(2, 1.0, 10.0)
    rng = np.random.Generator(np.random.PCG64(0))
    iters = 3000 if not (on_dev and persistent) else 100000
    nf.train_flow(rng, nf.elbo, flow, target, 10, max_iters=200, optimiser=nf.Adam(1e-3), ADbackend=nf.AutoNFCUDA(on_device=on_dev, chunk=1000 if not (on_dev and persistent) else 20000), show_progress=False)
    t0 = time.perf_counter()
    _, stats, _ = nf.train_flow(rng, nf.elbo, flow, target, 10, max_iters=iters, optimiser=nf.Adam(1e-3), ADbackend=nf.AutoNFCUDA(on_device=on_dev, chunk=1000 if not (on_dev and persistent) else 20000), show_progress=False)
    dt = time.perf_counter() - t0
    print("config 1 train_flow: on_device=%s persistent=%s  %.0f iterations/s  (%.1f us/iteration)  final loss %.4f" % (on_dev, persistent, iters / dt, 1e6 * dt / iters, stats[-1]["loss"]))

# coupling flow in the demo regime (reference example/demo_RealNVP.jl: 2-D banana, small MLPs): multi-launch loop vs CUDA-graph replay
for graph in (0, 1):
    os.environ["NFCUDA_TRAIN_GRAPH"] = str(graph)
    nf.seed(1)
    flow = nf.realnvp(nf.MvNormal(np.zeros(2)), [32, 32], 3, np.float32)
    target = nf.Banana(2, 1.0, 10.0)
    rng = np.random.Generator(np.random.PCG64(0))
    iters = 2000
    ad = nf.AutoNFCUDA(on_device=True, chunk=1000)
    nf.train_flow(rng, nf.elbo_batch, flow, target, 64, max_iters=100, optimiser=nf.Adam(1e-3), ADbackend=ad, show_progress=False)
    t0 = time.perf_counter()
    _, stats, _ = nf.train_flow(rng, nf.elbo_batch, flow, target, 64, max_iters=iters, optimiser=nf.Adam(1e-3), ADbackend=ad, show_progress=False)
    dt = time.perf_counter() - t0
    print("RealNVP d=2 [32,32] x6 couplings, batch 64, on-device Adam: graph=%d  %.0f iterations/s (%.1f us/iteration)  final loss %.4f"
          % (graph, iters / dt, 1e6 * dt / iters, stats[-1]["loss"]))
