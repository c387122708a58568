"""Seeded inputs and option sets on which the encoders are pinned to the reference (shared by the golden-vector
generator tests/golden/make_ref_golden.py and by tests/test_cpu_refpin.py)."""
import hashlib
import os

import numpy as np

from tests.conftest import GOLDEN
from tests.matrices import _csr_from_coo, poisson2d, random_structured, rmat, stencil27, sym_block_banded

XFORMS = ["none", "h", "v", "d", "ad", "br", "bc", "all", "h,d", "bc,v,ad", "br3{2,3},h{1}", "d{1},ad{2},v{1}"]


def ref_safe(rp, ci, va, n, m):
    """The reference's CSR iterator (Csr.hpp:256-373) mishandles an empty first row and never reaches end() behind
    trailing empty rows (it then reads past the arrays): give both rows an element."""
    rows = np.repeat(np.arange(n), np.diff(rp))
    e = [r for r in (0, n - 1) if rp[r + 1] == rp[r]]
    if not e:
        return rp, ci, va
    r = np.concatenate([rows, e])
    c = np.concatenate([ci, [min(x, m - 1) for x in e]])
    v = np.concatenate([va, [1.5] * len(e)])
    return _csr_from_coo(r, c, v, n, m)


def cases(full=False):
    """Yields (name, rowptr, colind, values, nrows, ncols, options)."""
    from oracle.pyoracle import OracleMatrix
    for name in ["demopatt", "test", "test2", "test3"]:   # the reference's bundled matrices, its test script's option sets
        M = OracleMatrix.from_mmf(os.path.join(GOLDEN, "matrices", name + ".mtx.sorted"))
        rp, ci, va = M.csr()
        for xf in XFORMS:
            for extra in ({}, {"spx.preproc.sampling": "none"}, {"spx.rt.nr_threads": 2}, {"spx.matrix.full_colind": "true"}):
                yield name, rp, ci, va, M.nrows, M.ncols, dict({"spx.preproc.xform": xf}, **extra)
    for name in ["symmetric", "symmetric-very-sparse"]:
        M = OracleMatrix.from_mmf(os.path.join(GOLDEN, "matrices", name + ".mtx.sorted"))
        rp, ci, va = M.csr()
        for xf in XFORMS:
            for extra in ({}, {"spx.preproc.sampling": "none"}, {"spx.rt.nr_threads": 2}):
                yield name, rp, ci, va, M.nrows, M.ncols, dict({"spx.preproc.xform": xf, "spx.matrix.symmetric": "true"}, **extra)
    for seed in range(12 if full else 4):
        rng = np.random.default_rng(seed)
        n, m = int(rng.integers(80, 600)), int(rng.integers(80, 600))
        rp, ci, va = ref_safe(*random_structured(rng, n, m), n, m)
        for xf in XFORMS:
            for extra in ({"spx.preproc.sampling": "none"}, {"spx.preproc.sampling": "none", "spx.rt.nr_threads": 3},
                          {"spx.preproc.sampling": "none", "spx.matrix.min_unit_size": 2, "spx.matrix.max_unit_size": 40},
                          {"spx.preproc.sampling": "none", "spx.matrix.split_blocks": "false"}):
                yield "rs%d" % seed, rp, ci, va, n, m, dict({"spx.preproc.xform": xf}, **extra)
    for seed in range(6 if full else 3):
        rng = np.random.default_rng(100 + seed)
        n = int(rng.integers(80, 500))
        rp, ci, va = ref_safe(*random_structured(rng, n, n, symmetric=True), n, n)
        for xf in XFORMS:
            for extra in ({"spx.preproc.sampling": "none"}, {"spx.preproc.sampling": "none", "spx.rt.nr_threads": 3}):
                yield "sym%d" % seed, rp, ci, va, n, n, dict({"spx.preproc.xform": xf, "spx.matrix.symmetric": "true"}, **extra)
    big = {"poisson200": poisson2d(200), "stencil27_30": stencil27(30), "rmat14": rmat(14), "symbb": sym_block_banded(20000, b=64)}
    for name, (rp, ci, va, n) in big.items():   # scaled-down BASELINE configs, default sampling
        for o in ({}, {"spx.rt.nr_threads": 4}, {"spx.preproc.xform": "br,bc"},
                  {"spx.preproc.sampling": "window", "spx.preproc.sampling.window_size": 500, "spx.preproc.sampling.nr_samples": 10}):
            yield name, rp, ci, va, n, n, o
        if name in ("poisson200", "symbb"):
            yield name, rp, ci, va, n, n, {"spx.matrix.symmetric": "true", "spx.rt.nr_threads": 2}


def key(name, opts):
    return name + "|" + ";".join("%s=%s" % kv for kv in sorted(opts.items()))


def digest(parts, sym):
    """Order-sensitive digest of an encoding: per partition row_start, nrows, id_map, ctl bytes, values (and dvalues)."""
    h = hashlib.sha256()
    for p in parts:
        h.update(np.asarray([p.row_start, p.nrows, len(p.ctl)], np.int64).tobytes())
        h.update(np.asarray(p.id_map, np.int64).tobytes())
        h.update(np.asarray(p.ctl, np.uint8).tobytes())
        h.update(np.asarray(p.values, np.float64).tobytes())
        if sym:
            h.update(np.asarray(p.dvalues, np.float64).tobytes())
    return h.hexdigest()
