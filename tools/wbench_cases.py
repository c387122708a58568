"""Scaled-down versions of the BASELINE configs shared by tools/wbench.py and tools/vbench.py (tuning aids)."""
from tests.matrices import poisson2d, rmat, stencil27, sym_block_banded

CASES = {
    "c2": (lambda: poisson2d(4096), {}),
    "c2s": (lambda: poisson2d(2048), {}),
    "c3s": (lambda: stencil27(160), {}),
    "c3b": (lambda: stencil27(128), {"spx.preproc.xform": "br,bc"}),
    "c4s": (lambda: sym_block_banded(1_000_000, b=1024), {"spx.matrix.symmetric": "true"}),
    "c4n": (lambda: sym_block_banded(1_000_000, b=1024), {}),
    "c5s": (lambda: rmat(22), {"spx.preproc.xform": "none"}),
}
