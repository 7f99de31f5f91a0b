"""Per-step device time of the C3 step next to SM clock / power / throttle reasons (diagnostic; run on the GPU box)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, pynvml
import nfload
nf = nfload.load(); K = nf._capi; lib = K.lib()
import bench
K.check(lib.nf_init(0))
flow = bench.make_theta(nf); tgt = nf.Funnel(64)
dev = torch.device("cuda", 0)
theta = torch.from_numpy(flow.theta).to(dev); z0 = torch.randn((1 << 20, 64), device=dev)
grad = torch.empty(flow.num_params, device=dev); val = C.c_double()
h, th = flow.handle(), tgt.handle()
pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
lib.nf_last_device_ms.restype = C.c_double
rows = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    t = time.perf_counter()
    K.check(lib.nf_elbo_value_and_grad_dev(h, th, theta.data_ptr(), 1 << 20, z0.data_ptr(), 0, -1.0, C.byref(val), grad.data_ptr()))
    wall = 1e3 * (time.perf_counter() - t)
    rows.append((i, wall, lib.nf_last_device_ms(h), pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM),
                 pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1e3,
                 pynvml.nvmlDeviceGetTemperature(hd, 0), hex(pynvml.nvmlDeviceGetCurrentClocksEventReasons(hd))))
for r in rows:
    print("%3d wall %7.2f dev %7.2f sm %4d mem %4d  %6.1f W  %2d C  %s" % r)
