"""Builds oracle/_ref/libcsxref_enc.so: the reference's own CSX encoder (SparseInternal, SparsePartition,
EncodingManager, Statistics, CsxManager, CtlBuilder ... headers and Encodings.cpp, Runtime.cpp, Statistics.cpp,
CtlBuilder.cpp, CsxUtil.cpp) compiled by g++ from the sources where they lie under /root/reference, against the
stand-in headers of oracle/refshim, plus the small driver oracle/ref_encoder.cpp.  The reference's build system
is not run (autotools + Boost + LLVM + libnuma are absent); no reference source is copied.  TEST INFRASTRUCTURE."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
SOURCES = ["Encodings.cpp", "Runtime.cpp", "Statistics.cpp", "CtlBuilder.cpp", "CsxUtil.cpp"]


def build(force=False):
    so = os.path.join(OUT, "libcsxref_enc.so")
    if not os.path.isdir(REF):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force and os.path.getmtime(so) > os.path.getmtime(os.path.join(HERE, "ref_encoder.cpp")):
        return so
    gen = os.path.join(OUT, "gen", "sparsex")
    os.makedirs(gen, exist_ok=True)
    # config.h as `configure` generates it for the default build (index int, value double)
    text = open(os.path.join(REF, "include", "sparsex", "config.h.in")).read()
    open(os.path.join(gen, "config.h"), "w").write(text.replace("@SPX_INDEX_TYPE@", "int").replace("@SPX_VALUE_TYPE@", "double"))
    flags = ["-std=c++14", "-O3", "-DNDEBUG", "-fPIC", "-w",  # -DNDEBUG -O3: the reference's release flags (m4check/ax_compilers.m4:260-261)
             "-I", os.path.join(HERE, "refshim"), "-I", os.path.join(OUT, "gen"),
             "-I", os.path.join(REF, "include")]
    objs = []
    for src in [os.path.join(REF, "src", "internals", s) for s in SOURCES] + [os.path.join(HERE, "ref_encoder.cpp")]:
        obj = os.path.join(OUT, "gen", os.path.basename(src) + ".o")
        subprocess.check_call(["g++"] + flags + ["-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-o", so] + objs)
    return so


if __name__ == "__main__":
    print("build_refenc: wrote", build(force="--force" in sys.argv))
