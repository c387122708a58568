"""Chunk-kernel decode logic on the CPU: the text of sparsex_b200/csrc/chunk_kernel.cuh (what nvcc compiles for
sm_100a) is compiled for the host and executed warp by warp on a lock-step fibre emulation of the warp
intrinsics (tests/emul/).  The product encoder and GPU layout builder feed it; results are compared with the
CSR input: decoded (row, column) per value bit-exact, y within 1e-12 (componentwise against |A||x|).
This is test infrastructure — the product has no CPU path."""
import numpy as np
import pytest

from tests.emul.run_emul import emul_spmv
from tests.matrices import random_structured, rmat, stencil27, sym_block_banded

TOL = 1e-12


def _check(rp, ci, va, n, m, opts, sym=False):
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, m)
    y, dr, dc, stats = emul_spmv(rp, ci, va, n, m, opts, 0.5, x)
    rows = np.repeat(np.arange(n), np.diff(rp))
    yref = np.zeros(n)
    np.add.at(yref, rows, va * x[ci])
    bound = np.zeros(n)
    np.add.at(bound, rows, np.abs(va * x[ci]))
    err = np.max(np.abs(y - 0.5 * yref) / (0.5 * bound + 1e-300))
    assert err <= TOL, (opts, err)
    if not sym:   # every stored value decodes to its CSR coordinates
        nnz = int(rp[-1])
        order = np.lexsort((dc[:nnz], dr[:nnz]))
        assert np.array_equal(dr[:nnz][order], rows) and np.array_equal(dc[:nnz][order], ci), opts
    return stats


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_emulated_chunk_kernel_all_unit_kinds(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(50, 400)), int(rng.integers(50, 400))
    rp, ci, va = random_structured(rng, n, m)
    for xf in ("all", "none", "h", "v", "d,ad", "br,bc", "h,v,d,ad,br,bc"):
        for sl in (0, 4, 12, 32):
            for fc in ("false", "true"):
                _check(rp, ci, va, n, m, {"spx.preproc.xform": xf, "spx.b200.slice": sl, "spx.matrix.full_colind": fc,
                                          "spx.preproc.sampling": "none", "spx.rt.nr_threads": 1 + seed})


@pytest.mark.parametrize("seed", [100, 101])
def test_emulated_chunk_kernel_symmetric(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(60, 300))
    rp, ci, va = random_structured(rng, n, n, symmetric=True)
    for xf in ("all", "none", "v", "d,ad", "br,bc", "h"):
        for sl in (0, 4, 24):
            for nt in (1, 3):
                _check(rp, ci, va, n, n, {"spx.preproc.xform": xf, "spx.b200.slice": sl, "spx.matrix.symmetric": "true",
                                          "spx.preproc.sampling": "none", "spx.rt.nr_threads": nt}, sym=True)


def test_emulated_gather_kernel_tile_shapes_and_diagonal_variant():
    """Kernel 1 (gather over the cross-row unit table): one and four rows per thread, the stride-1 diagonal
    instantiation the stencil configs run, carry-in descriptors across tiles, CSX-Sym transposed images."""
    from tests.matrices import poisson2d
    rp, ci, va, n = poisson2d(40, perturb=False)   # symmetric values: also used with spx.matrix.symmetric
    for rpt in (1, 4):
        for o in ({"spx.preproc.xform": "d", "spx.preproc.sampling": "none"}, {"spx.preproc.sampling": "none"},
                  {"spx.preproc.xform": "v,d,ad", "spx.preproc.sampling": "none", "spx.rt.nr_threads": 3},
                  {"spx.preproc.xform": "d,v", "spx.preproc.sampling": "none", "spx.matrix.symmetric": "true", "spx.rt.nr_threads": 2}):
            st = _check(rp, ci, va, n, n, dict(o, **{"spx.b200.rows_per_thread": rpt}), sym="spx.matrix.symmetric" in o)
            assert st[2] > 0   # table units present
    rng = np.random.default_rng(9)
    rp, ci, va = random_structured(rng, 700, 650)
    for rpt in (1, 4):
        _check(rp, ci, va, 700, 650, {"spx.preproc.xform": "v,d,ad", "spx.preproc.sampling": "none", "spx.b200.rows_per_thread": rpt})


def test_emulated_chunk_kernel_config_shapes():
    """Scaled-down BASELINE configs: block stencil, symmetric block-banded, R-MAT."""
    rp, ci, va, n = stencil27(14)
    st = _check(rp, ci, va, n, n, {"spx.preproc.xform": "br,bc", "spx.preproc.sampling": "none"})
    assert st[0] > 0 or st[12] > 0   # stream chunks, or everything in the block tables
    _check(rp, ci, va, n, n, {"spx.preproc.xform": "br,bc", "spx.preproc.sampling": "none", "spx.rt.nr_threads": 4})
    rp, ci, va, n = sym_block_banded(800, b=40)
    for o in ({}, {"spx.matrix.symmetric": "true"}, {"spx.matrix.symmetric": "true", "spx.rt.nr_threads": 3}):
        _check(rp, ci, va, n, n, dict(o, **{"spx.preproc.sampling": "none"}), sym="spx.matrix.symmetric" in o)
    rp, ci, va, n = rmat(11)
    for o in ({"spx.preproc.xform": "none"}, {"spx.preproc.xform": "none", "spx.b200.slice": 32},
              {"spx.preproc.xform": "none", "spx.rt.nr_threads": 5}):
        _check(rp, ci, va, n, n, o)


def _blocks(nbr, r, c, per_row, seed, ncols_b):
    from tests.matrices import _csr_from_coo
    rng = np.random.default_rng(seed)
    S = set()
    for I in range(nbr):
        for J in rng.choice(ncols_b, size=per_row, replace=False):
            for a in range(r):
                for b in range(c):
                    S.add((I * r + a, int(J) * c + b))
        S.add((I * r, int(rng.integers(ncols_b * c))))
    a = np.array(sorted(S))
    rp, ci, va = _csr_from_coo(a[:, 0], a[:, 1], rng.standard_normal(a.shape[0]), nbr * r, ncols_b * c)
    return rp, ci, va, nbr * r, ncols_b * c


def test_emulated_stream_kernel_shaped_instantiations(monkeypatch):
    """Block units that cannot live in the block tables stay with the stream kernel (here: tables switched off).  The
    pattern-set instantiations with compile-time block shapes (stream_kernel.cuh: SK_INSTANCES) are the ones the
    dispatcher picks for uniform blocks, and they multiply correctly (stats[3] = R*1000 + BC*100 + BRC*10)."""
    monkeypatch.setenv("CSXB_NO_BLOCK_TABLES", "1")
    want = {(3, 3, "bc"): 3300, (2, 2, "bc"): 2200, (4, 2, "bc"): 4200, (2, 4, "br"): 2040, (3, 8, "br"): 3040, (3, 3, "br"): 4000}
    for (r, c, xf), inst in want.items():
        rp, ci, va, n, m = _blocks(200, r, c, 3, 1, 150)
        st = _check(rp, ci, va, n, m, {"spx.preproc.xform": xf, "spx.preproc.sampling": "none"})
        assert st[3] == inst and st[12] == 0, ((r, c, xf), st[3])


def test_emulated_block_tables():
    """Aligned block units live in the block tables of the gather kernel (gpu_layout.hpp: BlockTable): own rows of
    block-column and block-row units, CSX-Sym images of block-column units; unaligned ones stay with the stream kernel."""
    for (r, c, xf) in ((3, 3, "bc"), (2, 2, "bc"), (6, 2, "bc"), (2, 4, "br"), (3, 8, "br"), (3, 3, "br,bc")):
        rp, ci, va, n, m = _blocks(200, r, c, 3, 1, 150)
        for nt in (1, 3):
            st = _check(rp, ci, va, n, m, {"spx.preproc.xform": xf, "spx.preproc.sampling": "none", "spx.rt.nr_threads": nt})
            assert st[12] > 0, ((r, c, xf), st)
    from tests.matrices import sym_block_banded
    rp, ci, va, n = sym_block_banded(1500, b=64)
    for xf in ("bc", "br", "all"):
        for nt in (1, 2, 5):
            st = _check(rp, ci, va, n, n, {"spx.matrix.symmetric": "true", "spx.preproc.xform": xf, "spx.preproc.sampling": "none",
                                          "spx.rt.nr_threads": nt}, sym=True)
            assert st[12] > 0, (xf, nt, st)
    # rows shifted by one: block-column units start at odd rows, no common sub-block -> stream kernel
    rp, ci, va, n, m = _blocks(100, 2, 2, 3, 2, 80)
    rp2 = np.concatenate([[0], rp]).astype(np.int32)
    st = _check(rp2, ci, va, n + 1, m, {"spx.preproc.xform": "bc", "spx.preproc.sampling": "none"})


def test_emulated_stream_kernel_long_rows_gaps_tall_blocks():
    """Rows longer than a chunk (fix-ups), runs of empty rows longer than the row window (gaps), tall block columns
    cut into several tasks, rounds of 32 units with the cursor carried across them."""
    from tests.matrices import _csr_from_coo
    nfix = 0
    for seed in range(2):
        r = np.random.default_rng(seed)
        n, m = 2000, 3200
        S = set()
        for row in (3, 700, 701, 1500):
            for c in r.choice(m, size=int(r.integers(2200, 3000)), replace=False):
                S.add((row, int(c)))
        for _ in range(200):
            S.add((int(r.integers(n)), int(r.integers(m))))
        for _ in range(6):
            r0, c0, h = int(r.integers(n - 40)), int(r.integers(m - 4)), int(r.integers(5, 30))
            for i in range(h):
                for j in range(3):
                    S.add((r0 + i, c0 + j))
        a = np.array(sorted(S))
        rp, ci, va = _csr_from_coo(a[:, 0], a[:, 1], r.standard_normal(a.shape[0]), n, m)
        for xf in ("none", "all", "br,bc", "h,bc"):
            for fc in ("false", "true"):
                st = _check(rp, ci, va, n, m, {"spx.preproc.xform": xf, "spx.matrix.full_colind": fc, "spx.preproc.sampling": "none",
                                              "spx.rt.nr_threads": 1 + seed * 2})
                assert st[5] > 0   # gaps present
                nfix += int(st[4])
    assert nfix > 0   # rows shared by chunks (fix-up entries) were met
