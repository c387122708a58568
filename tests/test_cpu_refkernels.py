"""CPU tests against the reference's own SpMV kernels (oracle/_ref: src/templates/*.c of SparseX assembled like
CsxJit does and compiled by gcc).  Pins the multiply half of the oracle and shows that the CSX streams
the engine emits are consumable by the reference's kernels unchanged."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN
from tests.matrices import poisson2d, random_structured, stencil27, sym_block_banded

XFORMS = ["none", "h", "v", "d", "ad", "br", "bc", "all", "bc,v,ad", "br3{2,3},h{1}", "d{1},ad{2},v{1}"]


def _ref():
    from oracle import refkernels
    if not refkernels.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refkernels


def _truth(rp, ci, va, x, n):
    rows = np.repeat(np.arange(n), np.diff(rp))
    y = np.zeros(n)
    np.add.at(y, rows, va * x[ci])
    return y


@pytest.mark.parametrize("name", ["demopatt", "test", "test3", "symmetric", "symmetric-very-sparse"])
def test_oracle_streams_run_on_reference_kernels(name):
    """Oracle-encoded CSX -> reference kernels == CSR product; and the oracle's own unit loops give the
    bit-identical vector (same operation order, no contraction)."""
    from oracle.pyoracle import OracleMatrix
    rk = _ref()
    M = OracleMatrix.from_mmf(os.path.join(GOLDEN, "matrices", name + ".mtx.sorted"))
    rp, ci, va = M.csr()
    x = np.random.default_rng(5).uniform(-1, 1, M.ncols)
    truth = _truth(rp, ci, va, x, M.nrows)
    for xf in XFORMS:
        for extra in ({}, {"spx.rt.nr_threads": 2}, {"spx.preproc.sampling": "none"}, {"spx.matrix.full_colind": "true"}):
            for sym in (("false", "true") if name.startswith("symmetric") else ("false",)):
                o = {"spx.preproc.xform": xf, "spx.matrix.symmetric": sym, "oracle.undefined_sampling": "break"}
                o.update(extra)
                M.tune(o)
                R = rk.Runner(M, full_colind=extra.get("spx.matrix.full_colind") == "true")
                y = R.spmv(0.5, x)
                assert np.abs(y - 0.5 * truth).max() <= 1e-12 * max(1.0, np.abs(truth).max()), (o, M.log)
                assert np.array_equal(y, M.spmv(0.5, x)), (o, M.log)


@pytest.mark.parametrize("seed", range(3))
def test_engine_streams_run_on_reference_kernels(seed):
    """CSX emitted by the engine's encoder (C-ABI) is fed to the reference's kernels unchanged."""
    from sparsex_b200 import CsxMatrix
    rk = _ref()
    rng = np.random.default_rng(40 + seed)
    for trial in range(3):
        sym = trial == 0
        n = int(rng.integers(20, 500))
        m = n if sym else int(rng.integers(20, 500))
        rp, ci, va = random_structured(rng, n, m, symmetric=sym)
        x = rng.uniform(-1, 1, m)
        truth = _truth(rp, ci, va, x, n)
        for xf in XFORMS:
            for extra in ({}, {"spx.rt.nr_threads": 3}, {"spx.preproc.sampling": "none"}):
                for s in (("true", "false") if sym else ("false",)):
                    o = {"spx.preproc.xform": xf, "spx.matrix.symmetric": s}
                    o.update(extra)
                    A = CsxMatrix.tune_csr(rp, ci, va, n, m, o)
                    parts = [A.partition(p) for p in range(A.nparts)]
                    R = rk.Runner(A, parts=parts, symmetric=(s == "true"))
                    y = R.spmv(1.0, x)
                    assert np.abs(y - truth).max() <= 1e-12 * max(1.0, np.abs(truth).max()), (o, parts[0].log)
                    A.close()


@pytest.mark.parametrize("gen,opts", [(lambda: poisson2d(96), {}), (lambda: stencil27(16), {"spx.preproc.xform": "br,bc"}),
                                      (lambda: sym_block_banded(800, b=16), {"spx.matrix.symmetric": "true"})])
def test_config_shapes_on_reference_kernels(gen, opts):
    from sparsex_b200 import CsxMatrix
    rk = _ref()
    rp, ci, va, n = gen()
    x = np.random.default_rng(1).uniform(-1, 1, n)
    truth = _truth(rp, ci, va, x, n)
    A = CsxMatrix.tune_csr(rp, ci, va, n, n, opts)
    R = rk.Runner(A, parts=[A.partition(p) for p in range(A.nparts)], symmetric=A.symmetric)
    y = R.spmv(1.0, x)
    assert np.abs(y - truth).max() <= 1e-12 * np.abs(truth).max()
    assert R.bench(1.0, x, 2) > 0
