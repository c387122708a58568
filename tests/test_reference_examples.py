"""The reference's own example programs (src/examples/*.c of SparseX) compile UNCHANGED against include/sparsex/*.h
and link against libsparsex_b200.so.  They are compiled from where they lie under /root/reference (never copied) into
tests/_refex/ (git-ignored; the binaries travel to the GPU box, the sources do not), and run there on the GPU."""
import os
import subprocess

import pytest

from tests.conftest import ROOT

REF = "/root/reference/src/examples"
OUT = os.path.join(ROOT, "tests", "_refex")
PROGRAMS = ["csr_example", "mmf_example", "advanced_example", "matrix_caching_example_p1", "matrix_caching_example_p2",
            "reordering_example"]


def build_examples():
    import sparsex_b200
    libdir = os.path.dirname(sparsex_b200.lib_path())
    os.makedirs(OUT, exist_ok=True)
    for prog in PROGRAMS:
        subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-w", os.path.join(REF, prog + ".c"), "-I", os.path.join(ROOT, "include"),
                               "-L", libdir, "-lsparsex_b200", "-lm", "-Wl,-rpath," + libdir, "-o", os.path.join(OUT, prog)])


def test_reference_examples_compile_unchanged():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    build_examples()
    for prog in PROGRAMS:
        assert os.path.exists(os.path.join(OUT, prog))


@pytest.mark.gpu
@pytest.mark.parametrize("prog,args", [("csr_example", []), ("mmf_example", ["tests/golden/matrices/demopatt.mtx.sorted"]),
                                       ("advanced_example", ["tests/golden/matrices/demopatt.mtx.sorted"]),
                                       ("reordering_example", ["tests/golden/matrices/symmetric.mtx.sorted"])])
def test_reference_examples_run(prog, args):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = os.path.join(OUT, prog)
    if not os.path.exists(exe):
        pytest.skip("tests/_refex was not built (the reference tree is only present in the build container)")
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
