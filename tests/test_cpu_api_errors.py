"""Argument checks, error codes and the error-handler hook of the drop-in API (host logic, no GPU needed).

Every case names the check of the reference it mirrors (src/api/matvec.c, src/api/error.c of SparseX): same error code
through the installed handler (spx_err_set_handler, error.c:58-70), same return value."""
import ctypes as C
import os

import numpy as np
import pytest

from sparsex_b200 import load_spx_api, SpxApi
from tests.conftest import GOLDEN

SPX_FAILURE, SPX_SUCCESS = -1, 0
ERR_ARG_INVALID, ERR_FILE, ERR_INPUT_MAT, ERR_DIM, OUT_OF_BOUNDS = 2, 3, 4, 9, 12
WARN_TUNING_OPT, WARN_RUNTIME_OPT = 23, 24
HANDLER = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_ulong, C.c_char_p, C.c_char_p)


class Recorder(object):
    def __init__(self, api):
        self.api, self.calls = api, []

        def handler(code, src, line, func, fmt):
            self.calls.append((code, (func or b"").decode(), (fmt or b"").decode()))
        self._cb = HANDLER(handler)
        api.L.spx_err_get_handler.restype = C.c_void_p
        self._old = api.L.spx_err_get_handler()
        api.spx_err_set_handler(C.cast(self._cb, C.c_void_p))

    def take(self):
        c, self.calls = self.calls, []
        return c

    def close(self):
        self.api.spx_err_set_handler(self._old)


@pytest.fixture()
def api_rec():
    api = load_spx_api()
    api.spx_init()
    rec = Recorder(api)
    yield api, rec
    rec.close()


def test_input_argument_checks(api_rec):
    api, rec = api_rec
    rp = np.array([0, 1, 2], np.int32)
    ci = np.array([0, 1], np.int32)
    va = np.array([1.0, 2.0])
    # matvec.c:179-197: dimensions first, then rowptr, colind, values
    assert not api.spx_input_load_csr(rp.ctypes.data, ci.ctypes.data, va.ctypes.data, -1, 2)
    assert rec.take() == [(ERR_ARG_INVALID, "spx_input_load_csr", "invalid matrix dimensions")]
    for args, msg in (((None, ci.ctypes.data, va.ctypes.data), "invalid rowptr argument"),
                      ((rp.ctypes.data, None, va.ctypes.data), "invalid colind argument"),
                      ((rp.ctypes.data, ci.ctypes.data, None), "invalid values argument")):
        assert not api.spx_input_load_csr(*args, 2, 2)
        assert rec.take() == [(ERR_ARG_INVALID, "spx_input_load_csr", msg)]
    inp = api.spx_input_load_csr(rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 2, 2)
    assert inp and rec.take() == []
    assert api.spx_input_destroy(inp) == SPX_SUCCESS
    # matvec.c:222-243: no file name / unreadable file -> SPX_ERR_FILE
    assert not api.spx_input_load_mmf(None)
    assert [c[0] for c in rec.take()] == [ERR_FILE]
    assert not api.spx_input_load_mmf(b"/nonexistent/matrix.mtx")
    assert [c[0] for c in rec.take()] == [ERR_FILE]
    # an unsorted bare file is refused by the reader ("indices are not sorted in MMF file", Mmf.hpp:262-266)
    assert not api.spx_input_load_mmf(os.path.join(GOLDEN, "matrices", "demopatt.mtx.unsorted").encode())
    calls = rec.take()
    assert [c[0] for c in calls] == [ERR_INPUT_MAT]
    # matvec.c:247-253, 262-266
    assert api.spx_input_destroy(None) == SPX_FAILURE
    assert rec.take() == [(ERR_ARG_INVALID, "spx_input_destroy", "invalid input handle")]
    assert not api.spx_mat_tune(None)
    assert rec.take() == [(ERR_ARG_INVALID, "spx_mat_tune", "invalid input matrix")]


def test_matrix_handle_checks(api_rec):
    api, rec = api_rec
    v = C.c_double()
    for fn, args in ((api.spx_mat_get_nrows, ()), (api.spx_mat_get_ncols, ()), (api.spx_mat_get_nnz, ()),
                     (api.spx_mat_destroy, ())):
        assert fn(None, *args) == SPX_FAILURE
        assert rec.take()[0][:1] == (ERR_ARG_INVALID,)
    assert not api.spx_mat_get_partition(None)
    assert not api.spx_mat_get_perm(None)
    assert [c[2] for c in rec.take()] == ["invalid matrix handle"] * 2
    assert api.spx_mat_get_entry(None, 0, 0, C.byref(v)) == SPX_FAILURE
    assert api.spx_mat_set_entry(None, 0, 0, 1.0) == SPX_FAILURE
    assert api.spx_mat_save(None, b"/tmp/x") == SPX_FAILURE
    assert [c[0] for c in rec.take()] == [ERR_ARG_INVALID] * 3
    assert not api.spx_mat_restore(b"/nonexistent/file.csx")
    assert [c[0] for c in rec.take()] == [ERR_FILE]
    # matvec.c:557-571: handles are checked in the order matrix, x, y
    part = api.spx_partition_csr(np.array([0, 1, 2], np.int32).ctypes.data, 2, 1)
    x = api.spx_vec_create(2, part)
    assert api.spx_matvec_mult(1.0, None, x, x) == SPX_FAILURE
    assert [c[0::2] for c in rec.take()] == [(ERR_ARG_INVALID, "invalid matrix handle")]
    assert api.spx_matvec_kernel(1.0, None, x, 0.5, x) == SPX_FAILURE
    assert [c[0::2] for c in rec.take()] == [(ERR_ARG_INVALID, "invalid matrix handle")]
    api.spx_vec_destroy(x)
    api.spx_partition_destroy(part)


def test_partition_and_vector_helpers(api_rec):
    api, rec = api_rec
    rp = np.array([0, 2, 4, 6, 8, 10, 12, 14, 16], np.int32)     # 8 rows of two elements
    assert not api.spx_partition_csr(None, 8, 2)
    assert not api.spx_partition_csr(rp.ctypes.data, 8, 0)
    assert [c[0] for c in rec.take()] == [ERR_ARG_INVALID] * 2
    part = api.spx_partition_csr(rp.ctypes.data, 8, 2)            # matvec.c:687-735: rows split by non-zeros
    rs, re_ = api.spx_partition_get_rs(part), api.spx_partition_get_re(part)
    assert (rs[0], re_[0], rs[1], re_[1]) == (0, 4, 4, 8)
    assert not api.spx_partition_get_rs(None) and not api.spx_partition_get_re(None)
    assert api.spx_partition_destroy(None) == SPX_FAILURE
    assert [c[0::2] for c in rec.take()] == [(ERR_ARG_INVALID, "invalid partition handle")] * 3
    # vectors live in host memory without a GPU and the BLAS-1 helpers still work (Vector.cpp:259-377)
    assert not api.spx_vec_create_random(8, None)                 # matvec.c:820-829: needs a partition
    assert [c[2] for c in rec.take()] == ["invalid partition handle"]
    a, b, c = (api.spx_vec_create(8, part) for _ in range(3))
    an, bn = np.arange(8.0) + 1, np.linspace(-1, 1, 8)
    SpxApi.as_numpy(a)[:] = an
    SpxApi.as_numpy(b)[:] = bn
    api.spx_vec_scale_add(a, b, c, 0.5)
    assert np.allclose(SpxApi.as_numpy(c), an + 0.5 * bn, rtol=0, atol=1e-15)
    api.spx_vec_sub(a, b, c)
    assert np.array_equal(SpxApi.as_numpy(c), an - bn)
    assert abs(api.spx_vec_mul(a, b) - float(an @ bn)) < 1e-12
    assert api.spx_vec_compare(a, a) == 0 and api.spx_vec_compare(a, b) < 0
    # spx_vec_reorder / inv_reorder (matvec.c:933-980): permuted[p[i]] = v[i] and back; an invalid permutation is an error
    perm = np.array([3, 0, 7, 1, 6, 2, 5, 4], np.int32)
    pp = perm.ctypes.data_as(C.POINTER(C.c_int))
    assert api.spx_vec_reorder(a, pp) == SPX_SUCCESS
    assert np.array_equal(SpxApi.as_numpy(a)[perm], an)
    assert api.spx_vec_inv_reorder(a, pp) == SPX_SUCCESS
    assert np.array_equal(SpxApi.as_numpy(a), an)
    assert api.spx_vec_reorder(a, None) == SPX_FAILURE and api.spx_vec_inv_reorder(a, None) == SPX_FAILURE
    assert rec.take() == [(ERR_ARG_INVALID, "spx_vec_reorder", "invalid permutation"),
                          (ERR_ARG_INVALID, "spx_vec_inv_reorder", "invalid permutation")]
    # user buffers: a NULL buffer and an unknown mode are refused (matvec.c:781-803)
    buf = np.zeros(8)
    tuned = C.c_void_p()
    assert not api.spx_vec_create_from_buff(None, C.byref(tuned), 8, part, 43)
    assert not api.spx_vec_create_from_buff(buf.ctypes.data, C.byref(tuned), 8, part, 7)
    assert not api.spx_vec_create_from_buff(buf.ctypes.data, C.byref(tuned), 8, None, 44)   # SPX_VEC_TUNE needs a partition
    assert [c[2] for c in rec.take()] == ["invalid buffer", "invalid vector mode", "invalid partition handle"]
    v = api.spx_vec_create_from_buff(buf.ctypes.data, C.byref(tuned), 8, None, 43)
    assert v and tuned.value == buf.ctypes.data
    api.spx_vec_destroy(v)
    for h in (a, b, c):
        api.spx_vec_destroy(h)
    api.spx_partition_destroy(part)


def test_option_warnings(api_rec):
    api, rec = api_rec
    api.spx_option_set(None, b"1")                                   # matvec.c:753-756 -> warning, nothing set
    assert [c[0] for c in rec.take()] == [WARN_TUNING_OPT]
    api.spx_option_set(b"spx.preproc.no_such_option", b"1")          # unknown mnemonic (Runtime.hpp:108-134)
    assert [c[0] for c in rec.take()] == [WARN_TUNING_OPT]
    api.spx_option_set(b"spx.rt.no_such_option", b"1")
    assert [c[0] for c in rec.take()] == [WARN_RUNTIME_OPT]
    api.spx_option_set(b"spx.matrix.symmetric", b"maybe")            # a value that does not parse
    assert [c[0] for c in rec.take()] == [WARN_TUNING_OPT]
    api.spx_option_set(b"spx.preproc.xform", b"all")
    api.spx_option_set(b"spx.rt.nr_gpus", b"2")                      # engine additions are accepted silently
    api.spx_option_set(b"spx.b200.devices", b"")
    assert rec.take() == []


def test_tune_without_a_gpu_fails_through_the_handler(api_rec):
    """No CPU fallback: without a usable GPU spx_mat_tune reports SPX_ERR_TUNED_MAT and returns SPX_INVALID_MAT."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    api, rec = api_rec
    rp = np.array([0, 1, 2], np.int32)
    ci = np.array([0, 1], np.int32)
    va = np.array([1.0, 2.0])
    inp = api.spx_input_load_csr(rp.ctypes.data, ci.ctypes.data, va.ctypes.data, 2, 2)
    assert not api.spx_mat_tune(inp)
    assert [c[0] for c in rec.take()] == [5]                        # SPX_ERR_TUNED_MAT
    api.spx_input_destroy(inp)
