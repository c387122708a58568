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
    assert st[0] > 0
    _check(rp, ci, va, n, n, {"spx.preproc.xform": "br,bc", "spx.preproc.sampling": "none", "spx.rt.nr_threads": 4})
    rp, ci, va, n = sym_block_banded(800, b=40)
    for o in ({}, {"spx.matrix.symmetric": "true"}, {"spx.matrix.symmetric": "true", "spx.rt.nr_threads": 3}):
        _check(rp, ci, va, n, n, dict(o, **{"spx.preproc.sampling": "none"}), sym="spx.matrix.symmetric" in o)
    rp, ci, va, n = rmat(11)
    for o in ({"spx.preproc.xform": "none"}, {"spx.preproc.xform": "none", "spx.b200.slice": 32},
              {"spx.preproc.xform": "none", "spx.rt.nr_threads": 5}):
        _check(rp, ci, va, n, n, o)
