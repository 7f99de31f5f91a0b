"""normalizingflows.jl_b200 -- B200-native (sm_100a) ELBO / log-likelihood value+gradient hot path of
NormalizingFlows.jl behind the reference's own interface.  See DESIGN.md and INTEGRATION.md.

The directory name contains a dot, so import it through `nfload.load()` at the repo root (or put the
repo root on sys.path and `importlib.import_module`-load it the way nfload does).
"""
from . import _capi
from . import dp
from ._capi import NFCudaError, NF_MMA_SIMT, NF_MMA_F16X3, NF_MMA_F16X1, LIB_PATH
from .api import *  # noqa: F401,F403
from .api import _prepare_gradient, _value_and_gradient, _Loss, _Target  # noqa: F401
