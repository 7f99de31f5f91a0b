#!/usr/bin/env python
"""Headline benchmark: ELBO + gradient samples/s of the RealNVP d=64 flow (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = one reverse-KL ELBO value+gradient (reference `_value_and_gradient` of `elbo_batch`,
src/optimize.jl:86) over a synthetic batch of 2^20 base draws per GPU through
`realnvp(q0, [256,256], 4)` = 8 AffineCouplings on `Funnel(64)`, Float32, random-init weights.

  value  : whole-job samples/s with Z0 and theta resident in HBM (device-pointer C-ABI call, plus the
           NCCL all-reduce of the P+1 gradient/ELBO sums when N > 1), timed with CUDA sync brackets.
  e2e    : same metric through the host-buffer C-ABI call a Julia `ccall` makes (theta + Z0 host->device
           from pinned memory, gradient + value device->host inside the timed region).
  roofline: dominant kernel class (tcgen05 256x256 GEMM), CUDA-event durations recorded on the library's
           stream during the timed region; algorithmic FLOPs = 2*n*256*256 per launch.
  cpu_baseline: the oracle port (torch CPU autograd restatement, NOT the Julia reference -- Julia is not
           installable here) timed on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

DIM, HDIMS, NLAYERS = 64, [256, 256], 4
BATCH_PER_GPU = 1 << 20
WORKLOAD = "realnvp_d64_8xAffineCoupling_mlp2x256_funnel64_reverseKL_elbo+grad"
FLOP_PER_SAMPLE = 3 * 2 * (8 * 2 * 81920)          # fwd + dgrad + wgrad MACs, no recompute


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops_sustained", 1442.4), d.get("bf16_tflops", 1683.6), d.get("hbm_gbs", 6460.9), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (the fields of the B200_PROFILING.md clocks line:
    clocks.sm, clocks.max.sm, clocks_event_reasons.*).  Read in-process through NVML every 100 ms -- a forked
    `nvidia-smi -lms` takes about a second to start on a fresh box and its start-up stalls kernel launches, which used to
    land inside short timed regions; nvidia-smi remains the fallback, started early and waited for."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.samples, self.stop_flag, self.nvml = index, None, [], False, None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self._sample_nvml()                      # first (slow) query happens here, before the warm-up
            threading.Thread(target=self._loop_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
            t0 = time.time()
            while not self.samples and time.time() - t0 < 10:     # wait out nvidia-smi's start-up
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        reasons = [name for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                                          ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                                          ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                                          ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)) if r & bit]
        self.samples.append((time.perf_counter(), float(sm), float(mx), reasons))

    def _loop_nvml(self):
        while not self.stop_flag:
            try:
                self._sample_nvml()
            except Exception:
                pass
            time.sleep(0.1)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            reasons = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                       if v.lower().startswith("active")]
            self.samples.append((time.perf_counter(), sm, mx, reasons))

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples that arrived inside [t_begin, t_end] (the timed region)."""
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        sel = [x for x in self.samples if t_begin is None or (t_begin <= x[0] <= t_end + 0.05)]
        sm = [x[1] for x in sel]; mx = [x[2] for x in sel]
        reasons = sorted({r for x in sel for r in x[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_theta(nf):
    nf.seed(123)
    flow = nf.realnvp(nf.MvNormal(np.zeros(DIM)), HDIMS, NLAYERS, np.float32)
    return flow


def run_side_config(args):
    """--config c1|c2|c4|c4b|c3x1: the other BASELINE.json configs (parity-test cases, reported for context; single GPU,
    device-resident inputs).  Not the driver's headline line."""
    import torch
    import nfload
    nf = nfload.load()
    K = nf._capi
    lib = K.lib()
    K.check(lib.nf_init(0))
    nf.seed(123)
    cfg = args.config
    if cfg == "c1":
        flow, tgt, n, name = nf.planarflow(nf.MvNormal(np.zeros(2)), 20, np.float32), nf.Banana(2, 1.0, 10.0), 10, "planar x20 d=2 Banana N=10"
    elif cfg == "c2":
        flow, tgt, n, name = nf.radialflow(nf.MvNormal(np.zeros(2)), 20, np.float32), nf.WarpedGauss(1.0, 0.12), 1 << 20, "radial x20 d=2 WarpedGauss N=2^20"
    elif cfg == "c2p":
        flow, tgt, n, name = nf.planarflow(nf.MvNormal(np.zeros(2)), 20, np.float32), nf.Banana(2, 1.0, 10.0), 1 << 20, "planar x20 d=2 Banana N=2^20"
    elif cfg in ("c4", "c4b"):
        nl = 4 if cfg == "c4" else 8
        flow, tgt, n, name = nf.nsf(nf.MvNormal(np.zeros(16)), [32, 32], 10, 5.0, nl, np.float32), nf.Cross(2.0, 0.15, 16), 1 << 20, "NSF d=16 K=10 B=5 [32,32] %d couplings Cross x8 N=2^20" % (2 * nl)
    elif cfg == "c4ll":
        flow, tgt, n, name = nf.nsf(nf.MvNormal(np.zeros(16)), [32, 32], 10, 5.0, 4, np.float32), None, 1 << 20, "NSF d=16 K=10 B=5 [32,32] 8 couplings, forward-KL loglikelihood value+grad on exact Cross x8 samples, N=2^20"
    elif cfg == "c3b":
        nf.seed(123)
        flow, tgt, n, name = nf.realnvp(nf.MvNormal(np.zeros(DIM)), HDIMS, 8, np.float32), nf.Funnel(DIM), 1 << 20, "RealNVP d=64 16 couplings (nlayers=8) 2x256 Funnel N=2^20"
    elif cfg in ("c5", "c5f64"):
        import math
        inner = nf.Funnel(2, -8.0, 5.0)
        dt_ = np.float64 if cfg == "c5f64" else np.float32
        flow, tgt, n, name = nf.hamiltonian_flow(inner, 15, 3, math.log(0.05), dt_), nf.JointTarget(inner), 1 << 20, "Hamiltonian flow 15 x (momentum affine + LeapFrog L=3) on Funnel(2,-8,5) + N(0,I) momentum, %s, N=2^20" % dt_.__name__
    elif cfg in ("c5x", "c5xf64"):
        # BASELINE config 5 at its stated size: 100-D synthetic logistic-regression posterior (n_data = 256), flow of the
        # reference demo (15 x (momentum affine + LeapFrog L=3)), joint target; warp-per-sample kernel (csrc/hmc_warp.cu)
        import math
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        rng = np.random.Generator(np.random.PCG64(7))
        X = rng.standard_normal((256, 100)) / 10.0
        yb = (rng.uniform(size=256) < 1.0 / (1.0 + np.exp(-X @ rng.standard_normal(100)))).astype(np.float64)
        inner = nf.LogReg(X, yb, 1.0)
        dt_ = np.float64 if cfg == "c5xf64" else np.float32
        flow, tgt, n, name = nf.hamiltonian_flow(inner, 15, 3, math.log(0.01), dt_), nf.JointTarget(inner), 1 << 16, "Hamiltonian flow 15 x (momentum affine + LeapFrog L=3) on a 100-D logistic-regression posterior (256 observations) + N(0,I) momentum, %s, N=2^16" % dt_.__name__
    elif cfg == "c3x1":
        flow, tgt, n, name = make_theta(nf), nf.Funnel(DIM), 1 << 20, "RealNVP C3 in NF_MMA_F16X1 (single fp16 pass, NOT parity grade)"
        flow.set_mma_mode(nf.NF_MMA_F16X1)
    else:
        raise SystemExit("unknown --config " + cfg)
    n = args.batch if args.batch != BATCH_PER_GPU else n
    dev = torch.device("cuda", 0)
    d = flow.dim
    tdt = torch.float64 if flow.paramtype == np.float64 else torch.float32
    theta_dev = torch.from_numpy(flow.theta).to(dev)
    z0 = torch.randn((n, d), device=dev, dtype=tdt)
    if cfg == "c4ll":   # exact samples of the product of 8 Cross(2, 0.15) blocks (reference example/targets/cross.jl:30-38)
        comp = torch.randint(0, 4, (n, d // 2), device=dev)
        mu, sg = 2.0, 0.15
        mx = torch.tensor([0.0, -mu, mu, 0.0], device=dev)[comp]; my = torch.tensor([mu, 1.0, 1.0, -mu], device=dev)[comp]
        sx = torch.tensor([sg, 1.0, 1.0, sg], device=dev)[comp]; sy = torch.tensor([1.0, sg, sg, 1.0], device=dev)[comp]
        z0 = torch.stack([mx + sx * z0[:, 0::2], my + sy * z0[:, 1::2]], dim=2).reshape(n, d).contiguous()
    grad = torch.empty(flow.num_params, device=dev, dtype=tdt)
    val = C.c_double()
    torch.cuda.synchronize()
    h = flow.handle()
    th = tgt.handle() if tgt is not None else None

    def step():
        if tgt is None:
            K.check(lib.nf_loglik_value_and_grad_dev(h, theta_dev.data_ptr(), n, z0.data_ptr(), -1.0, C.byref(val), grad.data_ptr()))
        else:
            K.check(lib.nf_elbo_value_and_grad_dev(h, th, theta_dev.data_ptr(), n, z0.data_ptr(), 0, -1.0, C.byref(val), grad.data_ptr()))
    for _ in range(max(args.warmup, 3)):
        step()
    lib.nf_launch_count(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        step()
        dev_ms += lib.nf_last_device_ms(h)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sustained, burst, hbm, src = peaks()
    v = n * args.steps / dt
    # profiled pass (CUDA events around each launch of the profiled kernel classes), outside the timed region
    K.check(lib.nf_profile_enable(h, 1))
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    K.check(lib.nf_profile_enable(h, 0))
    kbuf = C.create_string_buffer(4096)
    K.check(lib.nf_profile_keys(h, kbuf, 4096))
    prof = {}
    for key in [k for k in kbuf.value.decode().split(",") if k]:
        cnt, ms = C.c_int64(), C.c_double()
        K.check(lib.nf_profile_collect(h, key.encode(), C.byref(cnt), C.byref(ms)))
        prof[key] = {"launches": cnt.value, "total_ms": ms.value}
    if cfg in ("c4", "c4b", "c4ll") and prof.get("rqs_bwd", {}).get("launches"):
        # dominant spline kernel: reads the 3K-1 conditioner outputs of every (sample, transformed coordinate), writes the same
        # number of gradients, plus the coordinate, its incoming gradient (read + write) and the per-sample logdet gradient
        bytes_per_launch = rqs_bwd_bytes(n, d)
        avg_ms = prof["rqs_bwd"]["total_ms"] / prof["rqs_bwd"]["launches"]
        ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": RQS_BWD_DESC,
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "avg_launch_ms": avg_ms,
                "launches_per_step": prof["rqs_bwd"]["launches"] / args.steps, "peak_source": src + " copy bandwidth", "traffic": None}
    elif "ew_flow" in prof and prof["ew_flow"]["launches"]:
        avg_ms = prof["ew_flow"]["total_ms"] / prof["ew_flow"]["launches"]
        ach = n * 4 * d / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "ew_flow_kernel (fused forward + target + backward; one Z0 read is all the HBM traffic)",
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "avg_launch_ms": avg_ms,
                "note": "algorithmic bytes = 4*d per sample; the kernel is SFU/FP32-issue bound by design (tanh/log1p/exp per layer), see DESIGN.md"}
    else:
        roof = None
    print(json.dumps({"metric": "ELBO+grad samples/sec", "config": {"workload": name}, "value": v, "unit": "samples/s",
                      "ms_per_step": 1e3 * dt / args.steps, "device_ms_per_step": dev_ms / args.steps, "steps": args.steps,
                      "gpu_launches": int(lib.nf_launch_count(0)), "loss": val.value, "roofline": roof, "kernel_classes": prof}))


def run_reference(args):
    """--impl reference: the reference's CPU arithmetic for this path, restated (oracle port, kind 'port'), on all host
    cores, a bounded sample of the workload per step.  The Julia reference itself is not installable here (no Julia, no
    network; Bijectors / MonotonicSplines / Flux / Zygote are not vendored), so `julia_threads` is null."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import nf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = 1 << 14
    rng = np.random.Generator(np.random.PCG64(123))
    of = O.realnvp(DIM, HDIMS, NLAYERS, torch.float32, rng)
    ot = O.Funnel(DIM)
    xs = torch.from_numpy(O.synthetic_z0(n_s, DIM))
    th = of.theta()
    for _ in range(args.warmup):
        O.elbo_value_and_grad(of, ot, th, xs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.elbo_value_and_grad(of, ot, th, xs)
    dt = time.perf_counter() - t0
    val = n_s * args.steps / dt
    sample = "one ELBO value+gradient per step over 2^14 = %d of the 2^20 base draws (torch CPU autograd oracle, %d threads)" % (n_s, cores)
    print(json.dumps({
        "impl": "reference", "metric": "ELBO+grad samples/sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": n_s,
                   "note": "CPU restatement oracle, not the Julia reference; samples/s is per-sample work, so the 2^14-draw step "
                           "is a bounded sample of the 2^20-draw workload"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                         "julia_threads": None, "julia_threads_reason": "Julia is not installed in this image or on the GPU box"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline():
    import torch
    import nf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = 1 << 14
    rng = np.random.Generator(np.random.PCG64(123))
    of = O.realnvp(DIM, HDIMS, NLAYERS, torch.float32, rng)
    ot = O.Funnel(DIM)
    xs = torch.from_numpy(O.synthetic_z0(n_s, DIM))
    th = of.theta()
    O.elbo_value_and_grad(of, ot, th, xs[:1024])
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 10.0:
        O.elbo_value_and_grad(of, ot, th, xs)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n_s * reps / dt, "unit": "samples/s", "cores": cores, "kind": "port", "julia_threads": None,
            "julia_threads_reason": "Julia is not installed in this image or on the GPU box",
            "sample": "%d x ELBO value+gradient over %d of the 2^20 base draws (torch CPU autograd oracle; "
                      "CPU restatement, not the Julia reference)" % (reps, n_s)}


def collect_profile(K, lib, h):
    kbuf = C.create_string_buffer(4096)
    K.check(lib.nf_profile_keys(h, kbuf, 4096))
    prof = {}
    for key in [k for k in kbuf.value.decode().split(",") if k]:
        cnt, ms = C.c_int64(), C.c_double()
        K.check(lib.nf_profile_collect(h, key.encode(), C.byref(cnt), C.byref(ms)))
        prof[key] = {"launches": cnt.value, "total_ms": ms.value}
    return prof


def measured_traffic(kernel_key, n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this build (profiles/r2_traffic.json: bytes per sample at the captured N, scaled to this N); None when the
    capture does not cover the kernel -- never a guess."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        e = d[kernel_key]
        return {"bytes_per_launch": e["dram_bytes_per_sample"] * n_local, "source": e["source"]}
    except Exception:
        return None


def side_c4(nf, K, lib, torch, steps=5):
    """BASELINE config 4 (NSF d=16, K=10, B=5, [32,32], 8 couplings, Cross x 8, N = 2^20) measured in the same run so that the
    driver's record carries it: resident inputs, CUDA-event device time, dominant-kernel roofline."""
    nf.seed(123)
    n, d = 1 << 20, 16
    flow = nf.nsf(nf.MvNormal(np.zeros(d)), [32, 32], 10, 5.0, 4, np.float32)
    tgt = nf.Cross(2.0, 0.15, d)
    dev = torch.device("cuda", torch.cuda.current_device())
    theta_dev = torch.from_numpy(flow.theta).to(dev)
    z0 = torch.randn((n, d), device=dev, dtype=torch.float32)
    grad = torch.empty(flow.num_params, device=dev, dtype=torch.float32)
    val = C.c_double()
    torch.cuda.synchronize()
    h, th = flow.handle(), tgt.handle()

    def step():
        K.check(lib.nf_elbo_value_and_grad_dev(h, th, theta_dev.data_ptr(), n, z0.data_ptr(), 0, -1.0, C.byref(val), grad.data_ptr()))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(steps):
        step()
        dev_ms += lib.nf_last_device_ms(h)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    K.check(lib.nf_profile_enable(h, 1))
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    K.check(lib.nf_profile_enable(h, 0))
    prof = collect_profile(K, lib, h)
    _, _, hbm, src = peaks()
    roof = None
    if prof.get("rqs_bwd", {}).get("launches"):
        bytes_per_launch = rqs_bwd_bytes(n, d)
        avg_ms = prof["rqs_bwd"]["total_ms"] / prof["rqs_bwd"]["launches"]
        ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": RQS_BWD_DESC, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "avg_launch_ms": avg_ms, "peak_source": src + " copy bandwidth", "traffic": None,
                "note": "bytes = the kernels' own unfused traffic (logits in, gradient planes out); algorithmic bytes of the step are 64 B/sample"}
    return {"workload": "nsf_d16_K10_B5_mlp32x32_8xNeuralSplineCoupling_cross8_reverseKL_elbo+grad", "batch": n, "value": n * steps / dt,
            "unit": "samples/s", "ms_per_step": 1e3 * dt / steps, "device_ms_per_step": dev_ms / steps, "steps": steps, "loss": val.value,
            "roofline": roof, "kernel_classes": prof}


RQS_BWD_DESC = ("rqs_bwd_kernel<float,10> (spline backward, K = 10: estimate pass over every 64th tile + full pass + commit; logits in through "
                "the bulk-copy ring, gradient w.r.t. the conditioner output out as the fp16 hi/lo planes the GEMMs read)")


def rqs_bwd_bytes(n, d, K=10):
    """Bytes one spline-backward step moves for n samples (its own unfused traffic): per (sample, transformed coordinate) the
    3K-1 logits in and their share of the two fp16 planes out (rows padded to 64 columns), the coordinate, its incoming gradient,
    the outgoing gradient through the side buffer (write, read, write); per sample the logdet gradient."""
    c, P3 = d // 2, 3 * K - 1
    ld = ((c * P3 + 63) // 64) * 64
    return n * (c * (P3 * 4 + 5 * 4) + ld * 4) + n * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="base draws per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the side blocks (C4, strong scaling)")
    ap.add_argument("--config", default="c3", help="c3 (headline) | c1 | c2 | c2p | c3b | c3x1 | c4 | c4b | c4ll | c5 | c5f64 | c5x | c5xf64")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "c3":
        return run_side_config(args)

    import torch
    import torch.distributed as dist
    import nfload
    nf = nfload.load()
    K = nf._capi
    lib = K.lib()
    lib.nf_last_device_ms.restype = C.c_double

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    K.check(lib.nf_init(local))
    dev = torch.device("cuda", local)
    # torch.distributed carries only the launcher's plumbing (the 128-byte NCCL id, barriers, max-over-ranks of the timings);
    # the gradient all-reduce of the data plane is the ncclAllReduce inside libnfcuda (nf_elbo_value_and_grad_multi*)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(K.NF_UNIQUE_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(nf.dp.Comm.unique_id()), dtype=torch.uint8).clone()
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        comm = nf.dp.Comm.init_rank(world, rank, bytes(uid.cpu().numpy().tobytes()), local)
    else:
        comm = nf.dp.Comm.init_all([local])

    n_local = args.batch
    flow = make_theta(nf)
    target = nf.Funnel(DIM)
    P = flow.num_params
    h = flow.handle()
    th = target.handle()
    theta_dev = torch.from_numpy(flow.theta).to(dev)
    g = torch.Generator(device=dev); g.manual_seed(2024 + rank)
    z0_dev = torch.randn((n_local, DIM), device=dev, dtype=torch.float32, generator=g)   # resident synthetic base draws
    grad_dev = torch.empty(P, device=dev, dtype=torch.float32)
    val = C.c_double()
    n_total = n_local * world
    torch.cuda.synchronize()          # inputs were produced on torch's stream; the library runs on its own
    vp = C.c_void_p
    flows_a, tgts_a = (vp * 1)(h), (vp * 1)(th)

    def make_resident(n_tot, z_dev):
        th_a, z_a, g_a = (vp * 1)(theta_dev.data_ptr()), (vp * 1)(z_dev.data_ptr()), (vp * 1)(grad_dev.data_ptr())

        def step():
            K.check(lib.nf_elbo_value_and_grad_multi_dev(comm._h, flows_a, tgts_a, th_a, n_tot, z_a, 0, -1.0, C.byref(val), g_a))
            return val.value
        return step
    step_resident = make_resident(n_total, z0_dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step, steps):
        sync_all()
        t0 = time.perf_counter()
        dms = 0.0
        for _ in range(steps):
            step()
            dms += lib.nf_last_device_ms(h)      # CUDA events on the library stream around this step's kernels (no host time)
        sync_all()
        t1 = time.perf_counter()
        tm = torch.tensor([t1 - t0, dms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm[0].item()), float(tm[1].item()), t0, t1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started before the warm-up so that its own start-up is outside the timed region
    for _ in range(args.warmup):
        loss = step_resident()
    lib.nf_launch_count(1)
    dt, dev_ms, t0, t1 = timed(step_resident, args.steps)
    loss = val.value
    launches = lib.nf_launch_count(0)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = n_total * args.steps / dt
    # second pass over the same K steps with a CUDA-event pair around every profiled launch (per-kernel-class device time
    # for the roofline); kept out of the headline region because the extra event records perturb launch overlap
    K.check(lib.nf_profile_enable(h, 1))
    dt_prof, _, _, _ = timed(step_resident, args.steps)
    K.check(lib.nf_profile_enable(h, 0))
    prof = collect_profile(K, lib, h)
    sustained, burst, hbm, peak_src = peaks()

    # ---- roofline of the dominant kernel class (largest share of the profiled step) ----
    roofline = None
    timed_classes = {k: v for k, v in prof.items() if v["launches"]}
    if timed_classes:
        dom = max(timed_classes, key=lambda k: timed_classes[k]["total_ms"])
        avg_ms = prof[dom]["total_ms"] / prof[dom]["launches"]
        if dom == "fused_affine_fwd":
            # one launch = both conditioner networks of one coupling, forward: 2 x 81 920 MACs per sample
            flops = 2.0 * n_local * 2 * 81920
            kname = "fused_affine_fwd_kernel (both conditioners of a coupling, 3 Dense layers each, + coupling arithmetic; fp16x3 split)"
        elif dom.startswith("tc_gemm_n256_k256"):
            flops = 2.0 * n_local * 256 * 256
            kname = "tc_gemm_kernel<256> (256x256 Dense dgrad, fp16x3 split = 3 MMAs per useful MAC)"
        elif dom.startswith("tc_wgrad_m256_n256"):
            flops = 2.0 * n_local * 256 * 256
            kname = "tc_wgrad_kernel<256> (256x256 weight gradient, contraction over samples)"
        else:
            flops, kname = None, dom
        if flops is not None:
            ach = flops / (avg_ms * 1e-3) / 1e12
            tr = measured_traffic(dom, n_local)
            roofline = {"bound": "tensor", "kernel": kname, "class": dom, "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
                        "frac": ach / sustained, "frac_issued_mma": 3 * ach / sustained, "avg_launch_ms": avg_ms,
                        "launches_per_step": prof[dom]["launches"] / args.steps,
                        "share_of_step": prof[dom]["total_ms"] / (1e3 * dt_prof), "profiled_ms_per_step": 1e3 * dt_prof / args.steps,
                        "peak_source": peak_src + " bf16 dense, sustained",
                        "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr["source"] if tr else None}
    step_roof = {"achieved_tflops": value / world * FLOP_PER_SAMPLE / 1e12,
                 "frac_of_bf16_sustained": value / world * FLOP_PER_SAMPLE / 1e12 / sustained,
                 "frac_issued_mma": 3 * value / world * FLOP_PER_SAMPLE / 1e12 / sustained,
                 "achieved_hbm_algorithmic_gbs": value / world * 4 * DIM / 1e9}

    # ---- e2e: the host-buffer C-ABI call a Julia `ccall` makes (theta + this rank's Z0 rows host->device from pinned memory,
    #      all-reduced gradient + value device->host), and the training form (z0 = NULL: device Philox draws) ----
    z0_host_t = torch.empty((n_local, DIM), dtype=torch.float32).pin_memory()
    z0_host_t.copy_(z0_dev.cpu())
    z0_host = z0_host_t.numpy()
    theta_host = flow.theta
    grad_host = torch.empty(P, dtype=torch.float32).pin_memory().numpy()

    def step_e2e():
        K.check(lib.nf_elbo_value_and_grad_multi(comm._h, flows_a, tgts_a, K.ptr(theta_host), n_total, K.ptr(z0_host), 0, -1.0,
                                                 C.byref(val), K.ptr(grad_host)))

    it_seed = [1000]

    def step_e2e_train():
        it_seed[0] += 1
        K.check(lib.nf_elbo_value_and_grad_multi(comm._h, flows_a, tgts_a, K.ptr(theta_host), n_total, None, it_seed[0], -1.0,
                                                 C.byref(val), K.ptr(grad_host)))

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    dt2, _, _, _ = timed(step_e2e, e2e_steps)
    e2e = {"value": n_total * e2e_steps / dt2, "unit": "samples/s", "h2d_bytes_per_step": int(world * (n_local * DIM * 4 + P * 4)),
           "d2h_bytes_per_step": int(P * 4 + 8), "steps": e2e_steps,
           "call": "nf_elbo_value_and_grad_multi(theta_host, z0_host) -- host-supplied Z0, the parity form"}
    step_e2e_train()
    dt3, _, _, _ = timed(step_e2e_train, e2e_steps)
    e2e_train = {"value": n_total * e2e_steps / dt3, "unit": "samples/s", "h2d_bytes_per_step": int(world * P * 4),
                 "d2h_bytes_per_step": int(P * 4 + 8), "steps": e2e_steps,
                 "call": "nf_elbo_value_and_grad_multi(theta_host, z0 = NULL, seed) -- device Philox draws, the call "
                         "train_flow(elbo_batch, flow, logp, n) makes"}

    side = {}
    if not args.no_side:
        if world > 1 and BATCH_PER_GPU % world == 0:
            # strong scaling: ONE global batch of 2^20 split over the ranks (the all-reduce and launch latency now weigh
            # against 1/N of the compute)
            n_strong = BATCH_PER_GPU
            step_strong = make_resident(n_strong, z0_dev[: n_strong // world])
            for _ in range(3):
                step_strong()
            dts, dms, _, _ = timed(step_strong, args.steps)
            side["strong_scaling"] = {"global_batch": n_strong, "n_gpus": world, "value": n_strong * args.steps / dts, "unit": "samples/s",
                                      "ms_per_step": 1e3 * dts / args.steps, "device_ms_per_step": dms / args.steps, "steps": args.steps}
        if world == 1:
            try:
                side["c4"] = side_c4(nf, K, lib, torch)
            except Exception as e:   # noqa: a side block must not take the headline line down
                side["c4"] = {"error": str(e)}
            try:
                # the chunked fallback: same step with the workspace capped at 12 GB (the stash of 2^20 draws needs ~46 GB),
                # so the library walks the batch in sample chunks (forward + backward per chunk)
                K.check(lib.nf_flow_set_workspace_limit(h, 12 << 30))
                for _ in range(2):
                    step_resident()
                dtc, dmc, _, _ = timed(step_resident, 3)
                side["chunked_workspace_12GB"] = {"value": n_total * 3 / dtc, "unit": "samples/s", "ms_per_step": 1e3 * dtc / 3,
                                                  "device_ms_per_step": dmc / 3, "steps": 3}
                K.check(lib.nf_flow_set_workspace_limit(h, 64 << 30))
            except Exception as e:   # noqa
                side["chunked_workspace_12GB"] = {"error": str(e)}

    if rank == 0:
        out = {
            "metric": "ELBO+grad samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "device_ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": n_local, "global_batch": n_total, "params": int(P),
                       "parallelism": "dp%d (samples sharded, theta replicated, one ncclAllReduce of P+1 doubles inside libnfcuda)" % world,
                       "entry_point": "nf_elbo_value_and_grad_multi_dev (value) / nf_elbo_value_and_grad_multi (e2e)",
                       "mma_mode": "tcgen05 kind::f16, fp16 hi/lo split x3, fp32 accumulate (parity mode)",
                       "l2_policy": "inputs larger than L2: Z0 268 MB and ~40 GB of stashed activations stream through HBM every step"},
            "loss": loss, "e2e": e2e, "e2e_train": e2e_train, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "step_roofline": step_roof, "kernel_classes": prof, "side": side,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
