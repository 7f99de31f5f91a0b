"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/nfcuda.h declares, and fails loudly (error code + message, no crash, no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nfcuda.h")).read()
    return sorted(set(re.findall(r"NF_API\s+[\w\s\*]+?\b(nf_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol(nf):
    syms = header_symbols()
    assert len(syms) >= 30
    lib = C.CDLL(nf.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes signature table covers the whole header and nothing else
    assert sorted(nf._capi.SIGNATURES) == syms


def test_version(nf):
    assert nf._capi.lib().nf_version() == 100


def test_no_cpu_fallback(nf):
    """Without a usable sm_100 device every entry point returns an error; nothing computes on the CPU."""
    lib = nf._capi.lib()
    n = C.c_int(-1)
    status = lib.nf_device_count(C.byref(n))
    if status == 0 and n.value > 0:
        pytest.skip("a CUDA device is present; the loud-failure path is exercised on CPU-only machines")
    flow = nf.planarflow(nf.MvNormal(np.zeros(2)), 2, nf.Float32)
    with pytest.raises(nf.NFCudaError):
        nf.elbo(flow, nf.Banana(2, 1.0, 10.0), np.zeros((4, 2), np.float32))
    assert lib.nf_last_error()


def test_null_handles_are_rejected(nf):
    lib = nf._capi.lib()
    assert lib.nf_flow_num_params(None) == -1
    assert lib.nf_flow_dim(None) == -1
    assert lib.nf_flow_set_mma_mode(None, 0) < 0
    val = C.c_double()
    assert lib.nf_elbo_value_and_grad(None, None, None, 1, None, 0, 1.0, C.byref(val), None) < 0
    assert b"null" in lib.nf_last_error().lower()
    lib.nf_flow_destroy(None)
    lib.nf_target_destroy(None)
