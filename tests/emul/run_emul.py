"""ctypes front end of tests/emul/libchunk_emul.so (host emulation of the chunk kernel; test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "sparsex_b200", "csrc"), "encoder.o", "gpu_layout.o", "mmf.o"])
        subprocess.check_call(["make", "-s", "-C", HERE], stderr=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(HERE, "libchunk_emul.so"))
        vp, i64 = C.c_void_p, C.c_int64
        L.emul_spmv.restype = C.c_int
        L.emul_spmv.argtypes = [vp, vp, vp, i64, i64, C.c_char_p, C.c_double, vp, vp, vp, vp, vp, C.c_char_p, C.c_size_t]
        _LIB = L
    return _LIB


def emul_spmv(rp, ci, va, nrows, ncols, opts, alpha, x):
    """Returns (y, decoded rows, decoded cols, stats) of the emulated engine for the whole matrix."""
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    va = np.ascontiguousarray(va, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(nrows)
    nnz = int(rp[-1])
    dr = np.full(nnz + 1, -2, np.int32)
    dc = np.full(nnz + 1, -2, np.int32)
    stats = np.zeros(16, np.int64)
    err = C.create_string_buffer(1024)
    o = ";".join("%s=%s" % kv for kv in (opts or {}).items()).encode()
    rc = lib().emul_spmv(rp.ctypes.data, ci.ctypes.data, va.ctypes.data, nrows, ncols, o, alpha, x.ctypes.data,
                         y.ctypes.data, dr.ctypes.data, dc.ctypes.data, stats.ctypes.data, err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return y, dr, dc, stats
