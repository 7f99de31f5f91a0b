"""Builds libnfcuda.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python normalizingflows.jl_b200/build.py [--force]

The shared object lands next to this file (git-ignored, but it travels to the GPU box with gpurun).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnfcuda.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["elementwise_fwd_f32.cu", "elementwise_fwd_f64.cu", "elementwise_inv_f32.cu", "elementwise_inv_f64.cu",
           "elementwise_train_f32.cu", "elementwise_train_f64.cu", "tc_gemm.cu", "general.cu", "hmc_warp.cu", "nfcuda.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# no -split-compile: measured 6-20 % slower fused elementwise kernels on B200 (it blocks inlining across the split)


def _digest():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for fn in sorted(files):
            with open(os.path.join(root, fn), "rb") as fh:
                h.update(fn.encode()); h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "nfcuda.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, src + ".log")
        with open(log, "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
