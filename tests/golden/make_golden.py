"""Generates tests/golden/encodings.npz: CSX encodings of the reference's bundled matrices
(tests/golden/matrices/, copied from the reference's test/matrices/) under the option sets of
test/scripts/test-sparsex.sh.in, as emitted by the oracle (oracle/csx_oracle.cpp).

The reference itself ships no golden ctl streams and cannot be built in this image (Boost, LLVM 4-6,
libnuma are absent), so these vectors freeze the restatement after it was validated against
(i) the hand-derived known-answer vector of SURVEY.md Appendix C and (ii) the reference's own kernel
templates compiled by gcc (oracle/_ref), which decode every stream back to y = A*x.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import OracleMatrix  # noqa: E402

CASES = [
    ("demopatt.mtx.sorted", "spx.preproc.xform=none"),
    ("demopatt.mtx.sorted", "spx.preproc.xform=h"),
    ("demopatt.mtx.sorted", "spx.preproc.xform=v"),
    ("demopatt.mtx.sorted", "spx.preproc.xform=all"),
    ("demopatt.mtx.sorted", "spx.preproc.xform=all;spx.rt.nr_threads=2"),
    ("demopatt.mtx.sorted", "spx.preproc.xform=all;spx.rt.nr_threads=2;spx.preproc.sampling.nr_samples=1;spx.preproc.sampling.portion=0.4"),
    ("symmetric.mtx.sorted", "spx.preproc.xform=all;spx.matrix.symmetric=true"),
    ("symmetric.mtx.sorted", "spx.preproc.xform=all;spx.matrix.symmetric=true;spx.rt.nr_threads=2"),
    ("symmetric.mtx.sorted", "spx.preproc.xform=all;spx.preproc.sampling.nr_samples=2;spx.preproc.sampling.portion=0.4"),
    ("symmetric-very-sparse.mtx.sorted", "spx.preproc.xform=all;spx.matrix.symmetric=true"),
    ("test.mtx.sorted", "spx.preproc.xform=all;spx.preproc.sampling=none"),
    ("test2.mtx.sorted", "spx.preproc.xform=all;spx.preproc.sampling=none"),
    ("test3.mtx.sorted", "spx.preproc.xform=all;spx.preproc.sampling=none"),
]


def main():
    out = {}
    for fixture, optstr in CASES:
        opts = dict(kv.split("=") for kv in optstr.split(";") if kv)
        M = OracleMatrix.from_mmf(os.path.join(ROOT, "tests", "golden", "matrices", fixture)).tune(opts)
        key = fixture + "|" + optstr
        out[key + "|nparts"] = np.int64(len(M.parts))
        for p, P in enumerate(M.parts):
            out["%s|ctl%d" % (key, p)] = P.ctl
            out["%s|values%d" % (key, p)] = P.values
            out["%s|idmap%d" % (key, p)] = P.id_map
        print(key, M.log)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "encodings.npz"), **out)


if __name__ == "__main__":
    main()
