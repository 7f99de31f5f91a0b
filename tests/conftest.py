import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def nf():
    """The package over libnfcuda.so.  A fresh checkout has no built library (it is git-ignored): compile it first, exactly as
    `__graft_entry__.build()` does (nvcc cross-compiles sm_100a without a GPU; a few minutes)."""
    import nfload
    mod = nfload.load()
    if not os.path.exists(mod.LIB_PATH):
        nfload.build()
    return mod


@pytest.fixture(scope="session")
def gpu(nf):
    """Initialise device 0 through the C ABI; GPU tests fail (not skip) when the library is unusable."""
    nf._capi.check(nf._capi.lib().nf_init(0))
    return nf
