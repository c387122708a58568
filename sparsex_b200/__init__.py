"""sparsex_b200 — B200-native CSX SpMV engine behind the SparseX C API.

The product is ``libsparsex_b200.so`` (hand-written sm_100a CUDA kernels, C++
host encoder, C-ABI in ``include/csx_b200.h`` and the ``spx_*`` drop-in API in
``include/sparsex/``).  This package is the Python host-side mirror used by the
tests and by ``bench.py``: thin ctypes bindings, no compute of its own and no
CPU fallback — if the shared library is missing, importing the bindings fails.
"""
from .engine import (CsxMatrix, DeviceGroup, EngineError, PeerExchange, lib, lib_path,  # noqa: F401
                     load_spx_api, SpxApi)

__all__ = ["CsxMatrix", "DeviceGroup", "EngineError", "PeerExchange", "lib", "lib_path", "load_spx_api", "SpxApi"]
