"""Reverse Cuthill-McKee (spx_mat_tune(..., SPX_MAT_REORDER); reference Rcm.hpp:116-340 on boost::cuthill_mckee_ordering).

Boost is not available here; the pin against it is the sample output Boost publishes for its own example program
(test_boost_documentation_example).  Further checks:
* the product (flat arrays, sparsex_b200/csrc/rcm.cpp) equals the oracle (oracle/rcm_oracle.cpp, a structural
  restatement of the published BGL code) on seeded matrices: connected, disconnected, with isolated vertices,
  structurally non-symmetric, with long degree ties (std::sort on > 16 elements);
* a hand-derived known answer;
* perm is a bijection, P A P^T is what csxb_permute_csr returns, bandwidth shrinks on scrambled banded matrices.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from sparsex_b200 import engine
from oracle import pyoracle


def oracle_rcm(rowptr, colind, n, symmetric=0):
    L = C.CDLL(pyoracle.build())
    L.rcm_oracle_csr.restype = C.c_int
    L.rcm_oracle_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    perm = np.empty(n, dtype=np.int32)
    rc = L.rcm_oracle_csr(rowptr.ctypes.data, colind.ctypes.data, n, symmetric, perm.ctypes.data)
    return None if rc else perm


def csr(a):
    a = sp.csr_matrix(a)
    a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)


def scrambled_banded(n, half_bw, seed, density=0.7):
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for d in range(-half_bw, half_bw + 1):
        i = np.arange(max(0, -d), min(n, n - d))
        keep = rng.random(len(i)) < density if d else np.ones(len(i), bool)
        rows.append(i[keep]); cols.append(i[keep] + d)
    r, c = np.concatenate(rows), np.concatenate(cols)
    q = rng.permutation(n)
    a = sp.coo_matrix((rng.uniform(1, 2, len(r)), (q[r], q[c])), shape=(n, n))
    return sp.csr_matrix(a)


def cases():
    rng = np.random.default_rng(7)
    out = []
    out.append(("banded", scrambled_banded(500, 4, 1)))
    out.append(("banded_sparse", scrambled_banded(800, 9, 2, density=0.25)))
    # disconnected blocks + isolated vertices (diagonal-only rows)
    blocks = sp.block_diag([scrambled_banded(60, 2, 3), sp.identity(7), scrambled_banded(90, 3, 4), sp.identity(1)])
    q = rng.permutation(blocks.shape[0])
    out.append(("components", sp.csr_matrix(sp.csr_matrix(blocks)[q][:, q])))
    # structurally non-symmetric: edges from either triangle, some doubled
    out.append(("nonsym", sp.csr_matrix(sp.random(300, 300, 0.02, random_state=5, format="csr") + sp.identity(300))))
    # a star with many leaves of equal degree: sorts of > 16 tied elements
    n = 120
    star = sp.lil_matrix((n, n))
    star[0, 1:] = 1; star[1:, 0] = 1
    star[5, 6] = star[6, 5] = 1
    star.setdiag(2)
    out.append(("star", sp.csr_matrix(star)))
    # 2-D grid graph in natural order
    g = 17
    grid = sp.kronsum(sp.diags([1.0, 1.0], [-1, 1], shape=(g, g)), sp.diags([1.0, 1.0], [-1, 1], shape=(g, g))) + sp.identity(g * g)
    out.append(("grid", sp.csr_matrix(grid)))
    out.append(("random_sym", sp.csr_matrix(sp.random(400, 400, 0.01, random_state=9) + sp.random(400, 400, 0.01, random_state=9).T + sp.identity(400))))
    return out


@pytest.mark.parametrize("name,a", cases(), ids=[c[0] for c in cases()])
def test_product_equals_oracle(name, a):
    rp, ci, va = csr(a)
    n = a.shape[0]
    perm, bw = engine.rcm_csr(rp, ci, n)
    ref = oracle_rcm(rp, ci, n)
    assert perm is not None and ref is not None
    assert np.array_equal(perm, ref)
    assert np.array_equal(np.sort(perm), np.arange(n))          # a bijection
    # csxb_permute_csr == P A P^T
    orp, oci, ova = engine.permute_csr(rp, ci, va, perm)
    b = sp.csr_matrix((ova, oci, orp), shape=a.shape)
    p = sp.csr_matrix((np.ones(n), (perm, np.arange(n))), shape=(n, n))   # P[new, old] = 1
    want = sp.csr_matrix(p @ sp.csr_matrix(a) @ p.T)
    want.sort_indices()
    assert np.array_equal(want.indptr, orp) and np.array_equal(want.indices, oci) and np.array_equal(want.data, ova)
    # the logged bandwidths are those of the matrix before and after
    coo = sp.coo_matrix(a)
    off = coo.row != coo.col
    assert bw[0] == np.abs(coo.row[off] - coo.col[off]).max()
    assert bw[1] == np.abs(perm[coo.row[off]] - perm[coo.col[off]]).max()
    assert np.abs(b.tocoo().row - b.tocoo().col).max() == bw[1]


def test_bandwidth_shrinks():
    for seed in range(3):
        a = scrambled_banded(2000, 5, 10 + seed)
        rp, ci, _ = csr(a)
        perm, bw = engine.rcm_csr(rp, ci, a.shape[0])
        assert bw[1] <= 4 * 5 and bw[1] < bw[0] // 20


def test_known_answer():
    # path 3 - 0 - 2 - 1 plus the isolated vertex 4.  Component {0,1,2,3}: representative 0; BFS from 0 has levels
    # {0}, {3, 2}, {1}: ecc 2, spouse 1; from 1: {1}, {2}, {0}, {3}: ecc 3, spouse 3; 3 > 2, so r = 1, x = 3 and
    # from 3: ecc 3, not larger: start = 3.  Cuthill-McKee visit order: 3, 0, 2, 1, then 4; reversed: inv_perm =
    # [4, 1, 2, 0, 3], perm[old] = new -> [3, 1, 2, 4, 0].
    a = sp.lil_matrix((5, 5))
    for i, j in ((0, 3), (0, 2), (2, 1)):
        a[i, j] = a[j, i] = 1.0
    a.setdiag(1.0)
    rp, ci, _ = csr(a)
    perm, bw = engine.rcm_csr(rp, ci, 5)
    assert perm.tolist() == [3, 1, 2, 4, 0]
    assert oracle_rcm(rp, ci, 5).tolist() == [3, 1, 2, 4, 0]
    assert bw == (3, 1)


def test_boost_documentation_example():
    """The graph and the sample output of Boost's own example, libs/graph/example/cuthill_mckee_ordering.cpp as printed in
    the BGL documentation (doc/cuthill_mckee_ordering.html, "Sample Output"): original bandwidth 8; reverse Cuthill-McKee
    ordering starting at 6: 8 3 0 9 2 5 1 4 7 6; starting at 0: 9 1 4 6 7 2 8 5 3 0; without a starting vertex (the call
    Rcm.hpp:136 makes): 0 8 5 7 3 6 4 2 1 9; bandwidth 4 each time.  The printed sequence is inv_perm (new position -> old
    vertex).  Product and oracle must both reproduce it."""
    edges = [(0, 3), (0, 5), (1, 2), (1, 4), (1, 6), (1, 9), (2, 3), (2, 4), (3, 5), (3, 8), (4, 6), (5, 6), (5, 7), (6, 7)]
    eu = np.array([e[0] for e in edges], np.int32)
    ev = np.array([e[1] for e in edges], np.int32)
    O = C.CDLL(pyoracle.build())
    O.rcm_oracle_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    L = engine.lib()
    L.csxb_rcm_edges.restype = C.c_int
    L.csxb_rcm_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    published = {6: [8, 3, 0, 9, 2, 5, 1, 4, 7, 6], 0: [9, 1, 4, 6, 7, 2, 8, 5, 3, 0], -1: [0, 8, 5, 7, 3, 6, 4, 2, 1, 9]}
    for start, want in published.items():
        perm = np.zeros(10, np.int32)
        assert O.rcm_oracle_edges(eu.ctypes.data, ev.ctypes.data, len(edges), 10, start, perm.ctypes.data) == 0
        assert np.argsort(perm).tolist() == want
        perm2 = np.zeros(10, np.int32)
        bw = np.zeros(2, np.int64)
        assert L.csxb_rcm_edges(eu.ctypes.data, ev.ctypes.data, len(edges), 10, start, perm2.ctypes.data, bw.ctypes.data) == 0
        assert np.argsort(perm2).tolist() == want and bw.tolist() == [8, 4]
    # a named starting vertex only orders a connected graph
    eu2 = np.array([0, 2], np.int32); ev2 = np.array([1, 3], np.int32)
    assert L.csxb_rcm_edges(eu2.ctypes.data, ev2.ctypes.data, 2, 4, 0, perm2.ctypes.data, None) == 1
    assert L.csxb_rcm_edges(eu2.ctypes.data, ev2.ctypes.data, 2, 4, -1, perm2.ctypes.data, None) == 0


def test_no_offdiagonal_and_bad_arguments():
    rp, ci, _ = csr(sp.identity(6))
    assert engine.rcm_csr(rp, ci, 6) == (None, None)
    assert oracle_rcm(rp, ci, 6) is None
    perm = np.zeros(6, dtype=np.int32)
    L = engine.lib()
    assert L.csxb_rcm_csr(rp.ctypes.data, ci.ctypes.data, 6, 7, perm.ctypes.data, None) < 0   # not square
    bad = ci.copy(); bad[2] = 9
    assert L.csxb_rcm_csr(rp.ctypes.data, bad.ctypes.data, 6, 6, perm.ctypes.data, None) < 0   # column out of range


def test_container_keeps_permutation(tmp_path):
    a = scrambled_banded(300, 3, 21)
    rp, ci, va = csr(a)
    n = a.shape[0]
    perm, _ = engine.rcm_csr(rp, ci, n)
    orp, oci, ova = engine.permute_csr(rp, ci, va, perm)
    A = engine.CsxMatrix.tune_csr(orp, oci, ova, n, n, {})
    L = engine.lib()
    assert L.csxb_get_perm(A._h, None) == 0
    assert L.csxb_set_perm(A._h, perm.ctypes.data, n) == 0
    path = str(tmp_path / "m.csx").encode()
    assert L.csxb_save(A._h, path) == 0
    err = C.create_string_buffer(512)
    h = L.csxb_load(path, err, 512)
    assert h, err.value
    got = np.empty(n, dtype=np.int32)
    assert L.csxb_get_perm(h, got.ctypes.data) == n and np.array_equal(got, perm)
    L.csxb_destroy(h)
    # a container whose permutation is not a bijection is refused
    raw = bytearray(open(path, "rb").read())
    raw[-4:] = raw[-8:-4]
    open(path, "wb").write(raw)
    assert not L.csxb_load(path, err, 512) and b"permutation" in err.value
