"""Tuned-matrix container (spx_mat_save / spx_mat_restore) and single-entry access (spx_mat_get_entry /
spx_mat_set_entry) on the host copy of the CSX arrays — SURVEY.md section 8f rows 2 and 3.  The element search
walks the ctl stream (all unit kinds) and must find every non-zero of the CSR input and nothing else."""
import os

import numpy as np
import pytest

from sparsex_b200 import CsxMatrix
from tests.matrices import random_structured, sym_block_banded

XF = ["all", "none", "h", "v", "d,ad", "br,bc", "h,v,d,ad,br,bc"]


def _entries(rp, ci, va, n):
    rows = np.repeat(np.arange(n), np.diff(rp))
    return rows, ci, va


@pytest.mark.parametrize("seed", [0, 1])
def test_get_set_entry_every_unit_kind(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(60, 300)), int(rng.integers(60, 300))
    rp, ci, va = random_structured(rng, n, m)
    rows, cols, vals = _entries(rp, ci, va, n)
    present = set(zip(rows.tolist(), cols.tolist()))
    for xf in XF:
        for extra in ({}, {"spx.matrix.full_colind": "true"}, {"spx.rt.nr_threads": 3}):
            A = CsxMatrix.tune_csr(rp, ci, va, n, m, dict({"spx.preproc.xform": xf, "spx.preproc.sampling": "none"}, **extra))
            for r, c, v in zip(rows.tolist(), cols.tolist(), vals.tolist()):
                assert A.get_entry(r, c) == v, (xf, extra, r, c)
            for _ in range(300):   # absent entries are reported as such
                r, c = int(rng.integers(n)), int(rng.integers(m))
                if (r, c) not in present:
                    assert A.get_entry(r, c) is None
            k = int(rng.integers(len(rows)))
            assert A.set_entry(int(rows[k]), int(cols[k]), 123.5) and A.get_entry(int(rows[k]), int(cols[k])) == 123.5
            assert not A.set_entry(*next((r, c) for r in range(n) for c in range(m) if (r, c) not in present), 1.0)
            A.close()


def test_get_set_entry_symmetric():
    rng = np.random.default_rng(7)
    n = 200
    rp, ci, va = random_structured(rng, n, n, symmetric=True)
    rows, cols, vals = _entries(rp, ci, va, n)
    for xf in ("all", "none", "br,bc", "d,v"):
        for nt in (1, 3):
            A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.preproc.xform": xf, "spx.preproc.sampling": "none",
                                                      "spx.matrix.symmetric": "true", "spx.rt.nr_threads": nt})
            for r, c, v in zip(rows.tolist(), cols.tolist(), vals.tolist()):   # both triangles and the diagonal
                assert A.get_entry(r, c) == v, (xf, nt, r, c)
            assert A.set_entry(5, 5, -2.0) and A.get_entry(5, 5) == -2.0
            A.close()


def test_save_and_load_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    cases = [random_structured(rng, 300, 260) + (300, 260, {"spx.rt.nr_threads": 3, "spx.preproc.sampling": "none"}),
             sym_block_banded(400, b=20)[:3] + (1200, 1200, {"spx.matrix.symmetric": "true", "spx.rt.nr_threads": 2})]
    for rp, ci, va, n, m, opts in cases:
        A = CsxMatrix.tune_csr(rp, ci, va, n, m, opts)
        path = os.path.join(str(tmp_path), "m.csxb")
        A.save(path)
        B = CsxMatrix.load(path)
        assert (B.nrows, B.ncols, B.nnz, B.symmetric, B.nparts, B.nparts_total) == (A.nrows, A.ncols, A.nnz, A.symmetric, A.nparts, A.nparts_total)
        for p in range(A.nparts):
            P, Q = A.partition(p), B.partition(p)
            for f in ("values", "ctl", "id_map", "rows_info", "dvalues", "map_cpus", "map_pos"):
                assert np.array_equal(getattr(P, f), getattr(Q, f)), f
            assert (P.row_start, P.nrows, P.ncols, P.nnz, P.log) == (Q.row_start, Q.nrows, Q.ncols, Q.nnz, Q.log)
        A.close()
        B.close()
    with open(os.path.join(str(tmp_path), "junk"), "wb") as f:
        f.write(b"not a container")
    with pytest.raises(Exception):
        CsxMatrix.load(os.path.join(str(tmp_path), "junk"))
