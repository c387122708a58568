"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front end for oracle/libcsx_oracle.so (the CPU restatement of the
SparseX CSX encoder and SpMV unit semantics).  Imported by tests/, by
``__graft_entry__.smoke()`` and by the cpu_baseline / ``--impl reference``
legs of bench.py — never by the product package ``sparsex_b200``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libcsx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("csx_oracle.cpp", "oracle_capi.cpp", "rcm_oracle.cpp", "csx_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libcsx_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.csxo_from_csr.restype = C.c_void_p
        L.csxo_from_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.csxo_from_mmf.restype = C.c_void_p
        L.csxo_from_mmf.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.csxo_tune.restype = C.c_int
        L.csxo_tune.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.csxo_free.argtypes = [C.c_void_p]
        L.csxo_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_long)] * 3
        L.csxo_coo.argtypes = [C.c_void_p] * 4
        L.csxo_nparts.restype = C.c_int
        L.csxo_nparts.argtypes = [C.c_void_p]
        L.csxo_log.restype = C.c_char_p
        L.csxo_log.argtypes = [C.c_void_p]
        L.csxo_part_info.restype = C.c_long
        L.csxo_part_info.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.csxo_part_copy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.csxo_spmv.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        L.csxo_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.csxo_bench.restype = C.c_double
        L.csxo_bench.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def opts_str(opts):
    return ";".join("%s=%s" % (k, v) for k, v in (opts or {}).items()).encode()


class OracleError(RuntimeError):
    pass


class Part(object):
    """One CSX partition: the fields of csx_matrix_t (Csx.hpp:37-48)."""
    pass


class OracleMatrix(object):
    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep
        n = [C.c_long() for _ in range(3)]
        lib().csxo_dims(self._h, *[C.byref(v) for v in n])
        self.nrows, self.ncols, self.nnz = [v.value for v in n]
        self.parts = []

    @classmethod
    def from_csr(cls, rowptr, colind, values, nrows, ncols):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colind = np.ascontiguousarray(colind, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        h = lib().csxo_from_csr(rowptr.ctypes.data, colind.ctypes.data, values.ctypes.data, nrows, ncols)
        return cls(h)

    @classmethod
    def from_mmf(cls, path):
        err = C.create_string_buffer(512)
        h = lib().csxo_from_mmf(path.encode(), err, 512)
        if not h:
            raise OracleError(err.value.decode())
        return cls(h)

    def coo(self):
        r = np.empty(self.nnz, np.int32)
        c = np.empty(self.nnz, np.int32)
        v = np.empty(self.nnz, np.float64)
        lib().csxo_coo(self._h, r.ctypes.data, c.ctypes.data, v.ctypes.data)
        return r, c, v

    def csr(self):
        r, c, v = self.coo()
        rowptr = np.zeros(self.nrows + 1, np.int32)
        np.add.at(rowptr, r + 1, 1)
        return np.cumsum(rowptr).astype(np.int32), c, v

    def tune(self, opts=None):
        err = C.create_string_buffer(512)
        if lib().csxo_tune(self._h, opts_str(opts), err, 512) != 0:
            raise OracleError(err.value.decode())
        L = lib()
        self.parts = []
        self.symmetric = bool(opts and str(opts.get("spx.matrix.symmetric", "false")) == "true")
        for p in range(L.csxo_nparts(self._h)):
            P = Part()
            info = [L.csxo_part_info(self._h, p, w) for w in range(9)]
            P.nnz, P.nrows, P.ncols, P.row_start, P.ctl_size, P.row_jumps = info[:6]

            def grab(what, n, dt):
                a = np.empty(n, dt)
                if n:
                    L.csxo_part_copy(self._h, p, what, a.ctypes.data)
                return a
            P.values = grab(0, P.nnz, np.float64)
            P.ctl = grab(1, P.ctl_size, np.uint8)
            P.id_map = grab(2, info[6], np.int64)
            P.rows_info = grab(3, P.nrows * 3, np.int32).reshape(-1, 3)
            P.dvalues = grab(4, info[8], np.float64)
            P.map_cpus = grab(5, info[7], np.uint32)
            P.map_pos = grab(6, info[7], np.uint32)
            self.parts.append(P)
        self.log = L.csxo_log(self._h).decode()
        return self

    def spmv(self, alpha, x, beta=0.0, y=None):
        """y = alpha*A*x (+ beta*y if y given) with the reference's unit semantics."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        overwrite = y is None
        if y is None:
            y = np.zeros(self.nrows, np.float64)
        else:
            y = np.array(y, dtype=np.float64)
        lib().csxo_spmv(self._h, alpha, x.ctypes.data, beta, y.ctypes.data, int(overwrite))
        return y

    def decode(self, part):
        n = self.parts[part].nnz
        r = np.empty(n, np.int32)
        c = np.empty(n, np.int32)
        if n:
            lib().csxo_decode(self._h, part, r.ctypes.data, c.ctypes.data)
        return r, c

    def bench(self, alpha, x, loops):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.nrows, np.float64)
        return lib().csxo_bench(self._h, alpha, x.ctypes.data, y.ctypes.data, loops)

    def __del__(self):
        try:
            if self._h:
                lib().csxo_free(self._h)
                self._h = None
        except Exception:
            pass
