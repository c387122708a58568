"""A plain C program written against <sparsex/sparsex.h> builds against the engine (CPU check) and runs (GPU)."""
import os
import subprocess

import pytest

from tests.conftest import ROOT

EXE = os.path.join(ROOT, "tests", "c_api_example.bin")


def _build():
    import sparsex_b200
    libdir = os.path.dirname(sparsex_b200.lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", os.path.join(ROOT, "tests", "c_api_example.c"),
                           "-I", os.path.join(ROOT, "include"), "-L", libdir, "-lsparsex_b200", "-lm",
                           "-Wl,-rpath," + libdir, "-o", EXE])


def test_c_program_builds_against_the_drop_in_headers():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["48"], ["300"], ["1100"], ["128", "true"]])
def test_c_program_runs(args):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _build()
    out = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
