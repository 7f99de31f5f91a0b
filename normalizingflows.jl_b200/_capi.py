"""ctypes binding of libnfcuda (include/nfcuda.h).  No torch types cross this boundary.

The library is built in-tree by `build.py`; loading fails loudly when it is missing -- there is no
CPU fallback behind this module.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnfcuda.so")

NF_F32, NF_F64 = 0, 1
NF_PLANAR, NF_RADIAL, NF_AFFINE_COUPLING, NF_SPLINE_COUPLING, NF_SHIFT, NF_SCALE = 1, 2, 3, 4, 5, 6
NF_MOMENTUM_AFFINE, NF_LEAPFROG = 7, 8
NF_TARGET_BANANA, NF_TARGET_FUNNEL, NF_TARGET_WARPED_GAUSS, NF_TARGET_CROSS, NF_TARGET_DIAG_NORMAL = 1, 2, 3, 4, 5
NF_TARGET_LOGREG = 6
NF_MMA_SIMT, NF_MMA_F16X3, NF_MMA_F16X1 = 0, 1, 2
NF_UNIQUE_ID_BYTES = 128


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("mask_idx", C.POINTER(C.c_int)), ("n_mask", C.c_int),
                ("hdims", C.POINTER(C.c_int)), ("n_hidden", C.c_int), ("K", C.c_int), ("B", C.c_double),
                ("n_steps", C.c_int), ("score_target", C.c_void_p)]


class NFCudaError(RuntimeError):
    pass


# every symbol declared in include/nfcuda.h: name -> (restype, argtypes)
_vp, _i, _i64, _u64, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_size_t
SIGNATURES = {
    "nf_version": (_i, []),
    "nf_last_error": (C.c_char_p, []),
    "nf_init": (_i, [_i]),
    "nf_device_count": (_i, [C.POINTER(_i)]),
    "nf_synchronize": (_i, []),
    "nf_flow_create": (_i, [C.POINTER(_vp), C.POINTER(LayerDesc), _i, _i, _i]),
    "nf_flow_destroy": (None, [_vp]),
    "nf_flow_num_params": (_i64, [_vp]),
    "nf_flow_dim": (_i, [_vp]),
    "nf_flow_set_base": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "nf_flow_set_base_chol": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "nf_flow_set_mma_mode": (_i, [_vp, _i]),
    "nf_flow_set_workspace_limit": (_i, [_vp, _sz]),
    "nf_flow_param_offset": (_i64, [_vp, _i]),
    "nf_target_create": (_i, [C.POINTER(_vp), _i, _i, C.POINTER(_d), _i]),
    "nf_target_create_joint": (_i, [C.POINTER(_vp), _vp]),
    "nf_target_destroy": (None, [_vp]),
    "nf_target_logp": (_i, [_vp, _i, _vp, _i64, _vp, _vp]),
    "nf_elbo_value_and_grad": (_i, [_vp, _vp, _vp, _i64, _vp, _u64, _d, C.POINTER(_d), _vp]),
    "nf_elbo_value_and_grad_dev": (_i, [_vp, _vp, _vp, _i64, _vp, _u64, _d, C.POINTER(_d), _vp]),
    "nf_elbo_sums_dev": (_i, [_vp, _vp, _vp, _i64, _vp, _u64, _vp]),
    "nf_elbo_terms": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "nf_loglik_value_and_grad": (_i, [_vp, _vp, _i64, _vp, _d, C.POINTER(_d), _vp]),
    "nf_loglik_value_and_grad_dev": (_i, [_vp, _vp, _i64, _vp, _d, C.POINTER(_d), _vp]),
    "nf_train_elbo_adam": (_i, [_vp, _vp, _vp, _i64, _u64, _i, _i, _d, _d, _d, _d, _vp, _vp, C.POINTER(_d)]),
    "nf_forward": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "nf_inverse": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "nf_logpdf": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "nf_sample": (_i, [_vp, _vp, _i64, _u64, _vp]),
    "nf_base_sample": (_i, [_vp, _i64, _u64, _vp]),
    "nf_forward_stash": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "nf_backward": (_i, [_vp, _vp, _vp, _vp]),
    "nf_spline_bins": (_i, [_vp, _vp, _i64, _vp, C.POINTER(C.c_int32)]),
    "nf_rqs_bin_search": (_i, [_i, _vp, _vp, _i64, _i, C.POINTER(C.c_int32)]),
    "nf_tc_gemm_test": (_i, [_i64, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "nf_fused_schedule": (_i, [_i, _i, _i, _i, C.POINTER(C.c_ubyte), _i]),
    "nf_launch_count": (_i64, [_i]),
    "nf_set_option": (_i, [C.c_char_p, _i]),
    "nf_comm_init_all": (_i, [C.POINTER(_vp), _i, C.POINTER(_i)]),
    "nf_comm_unique_id": (_i, [_vp]),
    "nf_comm_init_rank": (_i, [C.POINTER(_vp), _i, _i, _vp, _i]),
    "nf_comm_size": (_i, [_vp]),
    "nf_comm_local_size": (_i, [_vp]),
    "nf_comm_local_device": (_i, [_vp, _i]),
    "nf_comm_local_rank": (_i, [_vp, _i]),
    "nf_comm_destroy": (None, [_vp]),
    "nf_elbo_value_and_grad_multi": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _vp, _u64, _d, C.POINTER(_d), _vp]),
    "nf_loglik_value_and_grad_multi": (_i, [_vp, C.POINTER(_vp), _vp, _i64, _vp, _d, C.POINTER(_d), _vp]),
    "nf_elbo_value_and_grad_multi_dev": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i64, C.POINTER(_vp), _u64, _d,
                                              C.POINTER(_d), C.POINTER(_vp)]),
    "nf_shard_range": (None, [_i64, _i, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "nf_last_device_ms": (_d, [_vp]),
    "nf_profile_enable": (_i, [_vp, _i]),
    "nf_profile_keys": (_i, [_vp, C.c_char_p, _i]),
    "nf_profile_collect": (_i, [_vp, C.c_char_p, C.POINTER(_i64), C.POINTER(_d)]),
}

_lib = None


def lib():
    """Load libnfcuda.so (once).  Raises NFCudaError when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NFCudaError(
                "libnfcuda.so is missing at %s: build it with `python normalizingflows.jl_b200/build.py` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        msg = lib().nf_last_error()
        raise NFCudaError("libnfcuda error %d: %s" % (status, msg.decode() if msg else "?"))


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
