"""One process per GPU on matrices too large for one host: every rank generates and tunes only the rows of its own
partition (csxb_tune_csr_slab).  The row ranges computed from the row lengths (tests/matrices.py: split_rows) are the
reference's nnz-balanced split, and a partition tuned from its slab equals the one tuned from the whole matrix."""
import numpy as np
import pytest

from sparsex_b200 import CsxMatrix
from tests import matrices as M


def _compare(gen, counts, n, opts, nparts, lower=None):
    rp, ci, va = gen(0, n)
    ranges = M.split_rows(counts, nparts, lower)
    o = dict(opts, **{"spx.rt.nr_threads": nparts})
    A = CsxMatrix.tune_csr(rp, ci, va, n, n, o)
    for p in range(nparts):
        P = A.partition(p)
        lo, cnt = ranges[p]
        rows_ref = len(P.dvalues) if o.get("spx.matrix.symmetric") == "true" else P.nrows
        assert (P.row_start, rows_ref) == (lo, cnt), (p, (P.row_start, rows_ref), (lo, cnt))
        srp, sci, sva = gen(lo, lo + cnt)
        Q = CsxMatrix.tune_csr_slab(srp, sci, sva, n, n, lo, p, o).partition(0)
        for f in ("ctl", "values", "id_map", "dvalues"):
            assert np.array_equal(getattr(P, f), getattr(Q, f)), (p, f)
        assert (P.row_start, P.nrows, P.nnz) == (Q.row_start, Q.nrows, Q.nnz)


def test_slab_tune_equals_whole_matrix_tune_stencils():
    g = 20
    _compare(lambda lo, hi: M.stencil_rows("s27", g, lo, hi), M.stencil_row_counts("s27", g), g ** 3, {}, 3)
    _compare(lambda lo, hi: M.stencil_rows("s27", g, lo, hi), M.stencil_row_counts("s27", g), g ** 3, {"spx.preproc.xform": "br,bc"}, 4)
    g = 90
    _compare(lambda lo, hi: M.stencil_rows("p2", g, lo, hi), M.stencil_row_counts("p2", g), g * g, {}, 5)


def test_slab_tune_equals_whole_matrix_tune_block_banded_and_symmetric():
    nb, b = 3000, 64
    n = nb * 3
    cnt = M.symbb_row_counts(nb, b)
    gen = lambda lo, hi: M.symbb_rows(nb, b, lo, hi)   # noqa: E731
    rp, ci, va = gen(0, n)
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert np.array_equal(np.diff(rp), cnt)
    # symmetric by construction of the hash values
    key = rows * n + ci
    tkey = ci.astype(np.int64) * n + rows
    assert np.array_equal(va[np.argsort(key)], va[np.argsort(tkey)])
    _compare(gen, cnt, n, {}, 3)
    lower = np.bincount(rows[ci < rows], minlength=n)
    for nparts in (1, 2, 4):
        _compare(gen, cnt, n, {"spx.matrix.symmetric": "true"}, nparts, lower)


def test_slab_tune_equals_whole_matrix_tune_rmat_blocks():
    scale = 14
    n = 1 << scale
    cnt = M.rmat_block_row_counts(scale, device="cpu")
    rp, ci, va = M.rmat_block_rows(scale, 0, n, device="cpu")
    assert np.array_equal(np.diff(rp), cnt)
    _compare(lambda lo, hi: M.rmat_block_rows(scale, lo, hi, device="cpu"), cnt, n, {"spx.preproc.xform": "none"}, 4)
