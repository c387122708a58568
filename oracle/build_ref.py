"""ORACLE — TEST INFRASTRUCTURE ONLY.

Builds oracle/_ref/libcsxref_gen.so from the reference's own SpMV kernel sources, read where they lie:
  /root/reference/src/templates/*.c                  (the 18 text templates the JIT assembles)
  /root/reference/include/sparsex/{types.h,config.h.in}, internals/{CtlUtil,Vector,Csx,Map}.hpp, cdecl.h, numa_util.h
The template/header text is embedded as data next to oracle/refgen.c (this repository's restatement of
CsxJit's assembly step).  At run time — also on the GPU box, where /root/reference does not exist —
oracle/refkernels.py asks the library for the translation unit of a partition's id_map and compiles it
with gcc -std=c99 -O3 -ffp-contract=off, i.e. gcc stands in for the Clang/LLVM JIT the reference needs.
The reference's own build system (autotools + Boost + LLVM 4-6 + libnuma) cannot run in this image.

Outputs go to oracle/_ref/ only (git-ignored; travels with gpurun).  Run: python oracle/build_ref.py
"""
import glob
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
HEADERS = ["sparsex/types.h", "sparsex/internals/CtlUtil.hpp", "sparsex/internals/Vector.hpp",
           "sparsex/internals/Csx.hpp", "sparsex/internals/Map.hpp", "sparsex/internals/cdecl.h",
           "sparsex/internals/numa_util.h"]


def c_string(text):
    out = []
    for line in text.split("\n"):
        out.append('"' + line.replace("\\", "\\\\").replace('"', '\\"') + '\\n"')
    return "\n".join(out)


def main():
    if not os.path.isdir(REF):
        print("build_ref: %s not present, keeping the prebuilt oracle/_ref" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        res = ["struct csxref_resource { const char *name; const char *text; };",
               "const struct csxref_resource csxref_templates[] = {"]
        for path in sorted(glob.glob(os.path.join(REF, "src", "templates", "*.c"))):
            res.append('{"%s",\n%s},' % (os.path.basename(path), c_string(open(path).read())))
        res.append("{0, 0}};")
        res.append("const struct csxref_resource csxref_headers[] = {")
        for h in HEADERS:
            res.append('{"%s",\n%s},' % (h, c_string(open(os.path.join(REF, "include", h)).read())))
        cfg = open(os.path.join(REF, "include", "sparsex", "config.h.in")).read()
        cfg = cfg.replace("@SPX_INDEX_TYPE@", "int").replace("@SPX_VALUE_TYPE@", "double")  # configure.ac:98,111
        res.append('{"sparsex/config.h",\n%s},' % c_string(cfg))
        res.append("{0, 0}};")
        rc = os.path.join(tmp, "resources.c")
        open(rc, "w").write("\n".join(res))
        subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-fPIC", "-shared", "-o", os.path.join(OUT, "libcsxref_gen.so"),
                               os.path.join(HERE, "refgen.c"), rc])
    print("build_ref: wrote", os.path.join(OUT, "libcsxref_gen.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
