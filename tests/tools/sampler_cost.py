"""How much does in-process NVML polling perturb a timed region?  (run on the GPU box; diagnostic only)"""
import os, sys, time, threading, ctypes as C
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
def step():
    K.check(lib.nf_elbo_value_and_grad_dev(h, th, theta.data_ptr(), 1 << 20, z0.data_ptr(), 0, -1.0, C.byref(val), grad.data_ptr()))
def timed(n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t) / n
for _ in range(3): step()
print("no sampler      ", [round(timed(), 2) for _ in range(3)])
pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
def cost(fn, name):
    ts = []
    for _ in range(5):
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
    print("  nvml %-28s %s ms" % (name, [round(x, 2) for x in ts]))
stop = False
def bg():
    while not stop: step()
th_bg = threading.Thread(target=bg); th_bg.start(); time.sleep(0.3)
cost(lambda: pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), "clock sm")
cost(lambda: pynvml.nvmlDeviceGetMaxClockInfo(hd, pynvml.NVML_CLOCK_SM), "max clock")
cost(lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(hd), "event reasons")
cost(lambda: pynvml.nvmlDeviceGetPowerUsage(hd), "power")
stop = True; th_bg.join()
for period, what in [(0.1, "all"), (0.5, "all"), (0.1, "clock"), (0.1, "reasons")]:
    run = True
    def poll():
        while run:
            if what in ("all", "clock"): pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM)
            if what in ("all", "reasons"): pynvml.nvmlDeviceGetCurrentClocksEventReasons(hd)
            time.sleep(period)
    t = threading.Thread(target=poll); t.start()
    print("poll %-8s every %.1fs" % (what, period), [round(timed(), 2) for _ in range(3)])
    run = False; t.join()
print("no sampler again", [round(timed(), 2) for _ in range(3)])
