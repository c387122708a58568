"""CPU tests: the oracle against the known-answer vector and the reference fixtures; the product encoder
(host side of spx_mat_tune, through the C-ABI) against the oracle; C-ABI symbol coverage."""
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT
from tests.matrices import poisson2d, random_structured, stencil27, sym_block_banded

XFORMS = ["none", "h", "v", "d", "ad", "br", "bc", "all", "h,d", "bc,v,ad", "br3{2,3},h{1}", "d{1},ad{2},v{1}"]


def _oracle():
    from oracle.pyoracle import OracleMatrix
    return OracleMatrix


def _csr_spmv(rp, ci, va, x, n):
    rows = np.repeat(np.arange(n), np.diff(rp))
    y = np.zeros(n)
    np.add.at(y, rows, va * x[ci])
    return y


def test_known_answer_demopatt_horizontal():
    """hand-derived ctl stream for demopatt, xform=h, one thread."""
    M = _oracle().from_mmf(os.path.join(GOLDEN, "matrices", "demopatt.mtx.sorted")).tune({"spx.preproc.xform": "h"})
    P = M.parts[0]
    want = bytes.fromhex("000400010105" "810107" "810400010503" "81050001020204" "8103000108" "810400010404"
                         "81020201" "800402010102" "800402" "800402010104")
    assert bytes(P.ctl) == want
    assert P.id_map.tolist() == [10001, 8, -1] and P.row_jumps == 0
    assert P.rows_info[:, 0].tolist() == [0, 6, 9, 15, 22, 27, 33, 37, 43, 46]
    assert P.rows_info[:, 1].tolist() == [0, 5, 6, 10, 15, 18, 22, 24, 29, 33]
    assert P.rows_info[:, 2].tolist() == [0] * 10
    r, c, v = M.coo()
    assert np.array_equal(P.values, v)


def test_golden_encodings():
    """Committed golden encodings (tests/golden/encodings.npz, made by tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN, "encodings.npz")
    g = np.load(path, allow_pickle=False)
    names = sorted({k.split("|")[0] + "|" + k.split("|")[1] for k in g.files})
    assert names
    from sparsex_b200 import CsxMatrix
    for key in names:
        fixture, optstr = key.split("|")
        opts = dict(kv.split("=") for kv in optstr.split(";") if kv)
        fpath = os.path.join(GOLDEN, "matrices", fixture)
        O = _oracle().from_mmf(fpath).tune(opts)
        A = CsxMatrix.tune_mmf(fpath, opts)
        nparts = int(g[key + "|nparts"])
        assert len(O.parts) == nparts == A.nparts
        for p in range(nparts):
            assert np.array_equal(O.parts[p].ctl, g["%s|ctl%d" % (key, p)])
            assert np.array_equal(O.parts[p].values, g["%s|values%d" % (key, p)])
            assert np.array_equal(O.parts[p].id_map, g["%s|idmap%d" % (key, p)])
            assert np.array_equal(A.partition(p).ctl, g["%s|ctl%d" % (key, p)])
            assert np.array_equal(A.partition(p).values, g["%s|values%d" % (key, p)])
        A.close()


@pytest.mark.parametrize("name", ["demopatt", "test", "test2", "test3", "symmetric", "symmetric-very-sparse"])
def test_oracle_decodes_to_input(name):
    """Every stream the oracle emits decodes back to the input matrix and multiplies like CSR
    (the reference's own check, test/src/CsxCheck.cpp:28-48, at 1e-12 instead of 1e-6)."""
    M = _oracle().from_mmf(os.path.join(GOLDEN, "matrices", name + ".mtx.sorted"))
    r, c, v = M.coo()
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, M.ncols)
    yref = np.zeros(M.nrows)
    np.add.at(yref, r, v * x[c])
    sym_ok = name.startswith("symmetric")
    for xf in XFORMS:
        for extra in ({}, {"spx.preproc.sampling": "none"}, {"spx.rt.nr_threads": 2}):
            for sym in (("true", "false") if sym_ok else ("false",)):
                o = {"spx.preproc.xform": xf, "spx.matrix.symmetric": sym, "oracle.undefined_sampling": "break"}
                o.update(extra)
                M.tune(o)
                y = M.spmv(0.5, x)
                assert np.abs(y - 0.5 * yref).max() <= 1e-12 * max(1.0, np.abs(yref).max())
                if sym == "false":
                    dr = np.concatenate([M.decode(i)[0] for i in range(len(M.parts))])
                    dc = np.concatenate([M.decode(i)[1] for i in range(len(M.parts))])
                    dv = np.concatenate([P.values for P in M.parts])
                    order = np.lexsort((dc, dr))
                    assert np.array_equal(dr[order], r) and np.array_equal(dc[order], c) and np.array_equal(dv[order], v)


def test_reference_script_option_sets_are_defined():
    """The 13 option sets of test/scripts/test-sparsex.sh.in never reach the reference's undefined
    sampling reads: the oracle in strict mode must accept them."""
    O = _oracle()
    demo = os.path.join(GOLDEN, "matrices", "demopatt.mtx.sorted")
    symm = os.path.join(GOLDEN, "matrices", "symmetric.mtx.sorted")
    vsp = os.path.join(GOLDEN, "matrices", "symmetric-very-sparse.mtx.sorted")
    mt = {"spx.rt.nr_threads": 2, "spx.rt.cpu_affinity": "0,1"}
    samp = {"spx.preproc.sampling": "portion", "spx.preproc.sampling.nr_samples": 2, "spx.preproc.sampling.portion": 0.4}
    samp1 = {"spx.preproc.sampling.nr_samples": 1, "spx.preproc.sampling.portion": 0.4}
    cases = [(demo, {"spx.preproc.xform": "none"}), (demo, {"spx.preproc.xform": "h"}), (demo, {"spx.preproc.xform": "v"}),
             (demo, {"spx.preproc.xform": "all"}), (symm, dict(samp, **{"spx.preproc.xform": "all", "spx.matrix.symmetric": "true"})),
             (demo, dict(mt, **{"spx.preproc.xform": "all"})), (demo, dict(mt, **dict(samp1, **{"spx.preproc.xform": "all"}))),
             (symm, {"spx.preproc.xform": "all", "spx.matrix.symmetric": "true"}),
             (vsp, {"spx.preproc.xform": "all", "spx.matrix.symmetric": "true"}),
             (symm, dict(samp, **{"spx.preproc.xform": "all"})),
             (symm, dict(mt, **{"spx.preproc.xform": "all", "spx.matrix.symmetric": "true"})),
             (symm, dict(mt, **dict(samp1, **{"spx.preproc.xform": "all", "spx.matrix.symmetric": "true"})))]
    for path, opts in cases:
        O.from_mmf(path).tune(opts)  # raises OracleError("undefined: ...") otherwise


def test_reference_script_failure_cases():
    """symmetric=true on a non-symmetric matrix and an unsorted MMF file must fail cleanly
    (test-sparsex.sh.in:207-224), in the oracle and in the engine."""
    from oracle.pyoracle import OracleError
    from sparsex_b200 import CsxMatrix, EngineError
    demo = os.path.join(GOLDEN, "matrices", "demopatt.mtx.sorted")
    with pytest.raises(OracleError):
        _oracle().from_mmf(demo).tune({"spx.matrix.symmetric": "true"})
    with pytest.raises(EngineError):
        CsxMatrix.tune_mmf(demo, {"spx.matrix.symmetric": "true"})
    uns = os.path.join(GOLDEN, "matrices", "demopatt.mtx.unsorted")
    with pytest.raises(OracleError):
        _oracle().from_mmf(uns)
    with pytest.raises(EngineError):
        CsxMatrix.tune_mmf(uns)


def _compare(O, A, ctx):
    assert len(O.parts) == A.nparts, ctx
    for i, P in enumerate(O.parts):
        Q = A.partition(i)
        for f in ("nnz", "nrows", "ncols", "row_start", "ctl_size", "row_jumps"):
            assert getattr(P, f) == getattr(Q, f), (ctx, i, f)
        assert np.array_equal(P.ctl, Q.ctl), (ctx, i, "ctl", O.log, Q.log)
        assert np.array_equal(P.values, Q.values), (ctx, i, "values")
        assert np.array_equal(P.id_map, Q.id_map), (ctx, i, "id_map")
        assert np.array_equal(P.rows_info.astype(np.int64), Q.rows_info), (ctx, i, "rows_info")
        assert np.array_equal(P.dvalues, Q.dvalues), (ctx, i, "dvalues")
        assert np.array_equal(P.map_cpus, Q.map_cpus) and np.array_equal(P.map_pos, Q.map_pos), (ctx, i, "map")


@pytest.mark.parametrize("seed", range(4))
def test_engine_encoder_matches_oracle_random(seed):
    """Host side of spx_mat_tune (C-ABI csxb_tune_csr) vs the oracle: bit-exact CSX arrays."""
    from sparsex_b200 import CsxMatrix
    rng = np.random.default_rng(seed)
    for trial in range(6):
        n = int(rng.integers(5, 400))
        sym = trial % 3 == 0
        m = n if sym else int(rng.integers(5, 400))
        rp, ci, va = random_structured(rng, n, m, symmetric=sym)
        O = _oracle().from_csr(rp, ci, va, n, m)
        for xf in XFORMS:
            extras = [{}, {"spx.preproc.sampling": "none"}, {"spx.rt.nr_threads": int(rng.integers(2, 6))},
                      {"spx.matrix.full_colind": "true", "spx.rt.nr_threads": 2}, {"spx.matrix.split_blocks": "false"},
                      {"spx.matrix.min_unit_size": 2, "spx.matrix.max_unit_size": int(rng.integers(8, 255)),
                       "spx.matrix.min_coverage": 0.01},
                      {"spx.preproc.sampling.nr_samples": int(rng.integers(1, 6)),
                       "spx.preproc.sampling.portion": float(rng.uniform(0.05, 0.9))}]
            for extra in extras:
                for s in (("true", "false") if sym else ("false",)):
                    o = {"spx.preproc.xform": xf, "spx.matrix.symmetric": s}
                    o.update(extra)
                    O.tune(dict(o, **{"oracle.undefined_sampling": "break"}))
                    A = CsxMatrix.tune_csr(rp, ci, va, n, m, o)
                    _compare(O, A, (seed, trial, o))
                    A.close()


@pytest.mark.parametrize("gen,opts", [
    (lambda: poisson2d(200), {}), (lambda: poisson2d(200), {"spx.rt.nr_threads": 8}),
    (lambda: poisson2d(128), {"spx.matrix.symmetric": "true", "spx.rt.nr_threads": 2}),
    (lambda: stencil27(28), {}), (lambda: stencil27(28), {"spx.preproc.xform": "br,bc"}),
    (lambda: sym_block_banded(3000, b=32), {"spx.matrix.symmetric": "true"}),
    (lambda: sym_block_banded(3000, b=32), {"spx.matrix.symmetric": "true", "spx.rt.nr_threads": 4}),
])
def test_engine_encoder_matches_oracle_configs(gen, opts):
    """Scaled-down versions of BASELINE.json's configs in the default (sampling) regime."""
    from sparsex_b200 import CsxMatrix
    rp, ci, va, n = gen()
    O = _oracle().from_csr(rp, ci, va, n, n).tune(opts)
    A = CsxMatrix.tune_csr(rp, ci, va, n, n, opts)
    _compare(O, A, opts)
    # partial ranges (one process per GPU encodes only its own partition) give the same partition
    nt = int(opts.get("spx.rt.nr_threads", 1))
    if nt > 1:
        B = CsxMatrix.tune_csr(rp, ci, va, n, n, opts, part_lo=nt - 1, part_hi=nt)
        assert B.nparts == 1 and B.part_lo == nt - 1
        assert np.array_equal(B.partition(0).ctl, O.parts[nt - 1].ctl)
        assert np.array_equal(B.partition(0).values, O.parts[nt - 1].values)
        B.close()
    A.close()


def test_c_abi_exports_every_declared_symbol():
    """libsparsex_b200.so loads on a CPU-only box and exports what include/*.h declare."""
    import ctypes
    from sparsex_b200 import lib
    L = lib()
    names = set()
    for hdr in ("csx_b200.h", os.path.join("sparsex", "matvec.h"), os.path.join("sparsex", "common.h"),
                os.path.join("sparsex", "error.h")):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"static inline[^{]*\{[^}]*\}", "", text)
        names |= set(re.findall(r"\b((?:spx|csxb)_\w+|err_handle|malloc_internal|free_internal)\s*\(", text))
    names -= {"spx_malloc", "spx_free", "spx_err_get_handler"} - {"spx_err_get_handler"}
    names -= {"spx_malloc", "spx_free"}
    assert len(names) > 70
    for n in sorted(names):
        assert hasattr(L, n), "missing export: " + n
    assert isinstance(L.csxb_last_error, ctypes._CFuncPtr)
