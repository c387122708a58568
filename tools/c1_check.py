"""BASELINE config 1: the reference's bundled MatrixMarket matrices, CSX tuned, y = alpha*A*x on one B200 through
spx_matvec_mult vs the reference's CPU path with one thread (its encoder output — the oracle is bit-identical to it —
multiplied by its own kernel templates compiled by gcc) vs a CSR product.  Prints one line per matrix / option set."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import OracleMatrix  # noqa: E402
from sparsex_b200 import CsxMatrix  # noqa: E402
from tests.conftest import GOLDEN  # noqa: E402

worst = 0.0
for name, sym in (("demopatt", False), ("test", False), ("symmetric", True), ("symmetric-very-sparse", True)):
    M = OracleMatrix.from_mmf(os.path.join(GOLDEN, "matrices", name + ".mtx.sorted"))
    rp, ci, va = M.csr()
    n, m = M.nrows, M.ncols
    x = np.random.default_rng(1).uniform(-1, 1, m)
    rows = np.repeat(np.arange(n), np.diff(rp))
    ycsr = 0.5 * np.bincount(rows, weights=va * x[ci], minlength=n)
    bound = 0.5 * np.bincount(rows, weights=np.abs(va * x[ci]), minlength=n) + 1e-300
    for xf in ("all", "none", "h,v,d,ad", "br,bc"):
        opts = {"spx.preproc.xform": xf, "spx.rt.nr_threads": 1}
        if sym:
            opts["spx.matrix.symmetric"] = "true"
        O = OracleMatrix.from_mmf(os.path.join(GOLDEN, "matrices", name + ".mtx.sorted")).tune(opts)
        ycpu = O.spmv(0.5, x)
        kind = "oracle interpreter"
        try:
            from oracle import refkernels
            if refkernels.available():
                ycpu = refkernels.Runner(O).spmv(0.5, x)
                kind = "reference kernels (gcc)"
        except Exception:
            pass
        A = CsxMatrix.tune_csr(rp, ci, va, n, m, opts).upload(0)
        same_ctl = all(np.array_equal(A.partition(p).ctl, O.parts[p].ctl) for p in range(A.nparts))
        ygpu = np.zeros(n)
        A.spmv_host(0.5, x, ygpu)
        e_cpu = float(np.max(np.abs(ygpu - ycpu) / bound))
        e_csr = float(np.max(np.abs(ygpu - ycsr) / bound))
        worst = max(worst, e_cpu, e_csr)
        print("%-22s %dx%d nnz %d  xform %-9s [%s]  ctl identical %s  |gpu-cpu|/(|A||x|) %.2e  |gpu-csr|/(|A||x|) %.2e  (cpu: %s, 1 thread)"
              % (name, n, m, len(va), xf, O.log.strip()[:30], same_ctl, e_cpu, e_csr, kind))
        A.close()
print("worst componentwise relative difference %.2e (tolerance 1e-12)" % worst)
sys.exit(0 if worst <= 1e-12 else 1)
