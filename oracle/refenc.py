"""ctypes front end of oracle/_ref/libcsxref_enc.so — the reference's own encoder compiled by g++
(oracle/build_refenc.py).  TEST INFRASTRUCTURE: used to pin oracle/csx_oracle.cpp and to generate
tests/golden/ref_encodings.npz in the container that has /root/reference."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.exists(os.path.join(HERE, "_ref", "libcsxref_enc.so"))


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libcsxref_enc.so"))
        L.refenc_tune.restype = C.c_void_p
        L.refenc_tune.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.refenc_nparts.argtypes = [C.c_void_p]
        L.refenc_info.restype = C.c_long
        L.refenc_info.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refenc_copy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.refenc_log.restype = C.c_char_p
        L.refenc_log.argtypes = [C.c_void_p, C.c_int]
        L.refenc_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class RefPart(object):
    pass


def tune(rowptr, colind, values, nrows, ncols, opts=None):
    """Encodes a zero-based CSR matrix with the reference encoder; returns the list of partitions
    (csx_matrix_t fields: ctl, values, id_map, nnz, nrows, ncols, row_start, row_jumps; dvalues for CSX-Sym)."""
    L = lib()
    rp = np.ascontiguousarray(rowptr, np.int32)
    # one element of slack: CSR::iterator::operator* (Csr.hpp:362) is evaluated once at the end position by
    # SparsePartition::SetElems (SparsePartition.hpp:520) before the end test
    ci = np.concatenate([np.asarray(colind, np.int32), np.zeros(1, np.int32)])
    va = np.concatenate([np.asarray(values, np.float64), np.zeros(1, np.float64)])
    o = ";".join("%s=%s" % (k, v) for k, v in (opts or {}).items()).encode()
    h = L.refenc_tune(rp.ctypes.data, ci.ctypes.data, va.ctypes.data, nrows, ncols, o)
    parts = []
    for p in range(L.refenc_nparts(h)):
        P = RefPart()
        P.nnz, P.nrows, P.ncols, P.row_start, ctl_size, P.row_jumps, idl, dvl = [L.refenc_info(h, p, w) for w in range(8)]

        def grab(what, n, dt):
            a = np.empty(n, dt)
            if n:
                L.refenc_copy(h, p, what, a.ctypes.data)
            return a
        P.values = grab(0, P.nnz, np.float64)
        P.ctl = grab(1, ctl_size, np.uint8)
        P.id_map = grab(2, idl, np.int64)
        P.dvalues = grab(3, dvl, np.float64)
        P.log = L.refenc_log(h, p).decode()
        parts.append(P)
    L.refenc_free(h)
    return parts
