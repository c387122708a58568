"""Generates tests/golden/ref_encodings.json: digests of the CSX encodings (ctl, values, id_map, partition rows,
dvalues) that the REFERENCE's own encoder produces for the seeded cases of tests/refpin_cases.py.  Runs only where
/root/reference exists (this builds oracle/_ref/libcsxref_enc.so from the reference sources, oracle/build_refenc.py).
    python tests/golden/make_ref_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_refenc, refenc  # noqa: E402
from tests import refpin_cases as rc  # noqa: E402

if __name__ == "__main__":
    assert build_refenc.build(), "needs /root/reference"
    out = {}
    for name, rp, ci, va, n, m, opts in rc.cases(full=True):
        parts = refenc.tune(rp, ci, va, n, m, opts)
        out[rc.key(name, opts)] = rc.digest(parts, str(opts.get("spx.matrix.symmetric")) == "true")
    path = os.path.join(ROOT, "tests", "golden", "ref_encodings.json")
    json.dump(out, open(path, "w"), indent=0, sort_keys=True)
    print("wrote %d reference digests to %s" % (len(out), path))
