"""Copies the round's bench lines from gpurun_out/ into profiles/ and rewrites profiles/r02_summary.txt from every
profiles/r02_bench_*.json (tracked evidence; gpurun_out/ is scratch).  A new line measured with --no-cpu-baseline keeps the
CPU arm of the line it replaces (same matrix, same kind of box, earlier in the round) with a note saying so.
Usage: python tools/collect_profiles.py"""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2_bench_*_n*.json"))):
    lines = [ln for ln in open(f) if ln.startswith("{")]
    if not lines:
        continue
    d = json.loads(lines[0])
    m = re.match(r"r2_bench_(.+)_n(\d+)(_\w+)?\.json", os.path.basename(f))
    out = os.path.join(ROOT, "profiles", "r02_bench_%s_n%s%s.json" % (m.group(1), m.group(2), m.group(3) or ""))
    if not d.get("cpu_baseline") and os.path.exists(out):
        old = json.load(open(out))
        if old.get("cpu_baseline"):
            d["cpu_baseline"] = dict(old["cpu_baseline"], note="carried over from the earlier run of this round on the same matrix "
                                     "(this line was measured with --no-cpu-baseline)")
    json.dump(d, open(out, "w"), indent=1)
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*_n*.json"))):
    d = json.load(open(f))
    m = re.match(r"r02_bench_(.+)_n(\d+)(_\w+)?\.json", os.path.basename(f))
    r = d["roofline"]
    rows.append((d["config"]["workload"] + (" (--exchange nccl)" if (m.group(3) or "") == "_nccl" else ""), d["n_gpus"], d["value"], d["ms_per_step"], r["frac"], r.get("kernel_only_ms"), d["e2e"]["value"],
                 (d.get("cpu_baseline") or {}).get("value"), d["detail"]["encoding_rank0"], d["detail"]["checks_vs_csr"]))
with open(os.path.join(ROOT, "profiles", "r02_summary.txt"), "w") as fo:
    fo.write("# bench lines of round 2 (profiles/r02_bench_<workload>_n<N>.json): GFLOP/s, ms per SpMV step, fraction of the measured HBM peak\n"
             "# (N > 1: of the step incl. exchange; kernel_only = the rank's kernels alone), end-to-end GFLOP/s through host buffers,\n"
             "# CPU reference arm on the same matrix (N = 1), encoding of rank 0, max error against CSR on sampled rows (inside the bench run)\n"
             "# The N = 1 lines of c3, c3b and c4 were measured after the block-table and host-path changes at the end of the round;\n"
             "# the N > 1 lines and c2 before them (c4 at N > 1 runs the four-rows-per-thread block-table kernel of that time).\n")
    for w, n, v, ms, fr, ko, e2e, cpu, enc, chk in sorted(rows):
        errs = [x for x in (chk.get("device_path_max_err"), chk.get("exchange_own_rows_max_err"), chk.get("host_buffer_path_max_err")) if x is not None]
        fo.write("%-60s N=%d %9.1f GFLOP/s %8.4f ms  frac %.3f  kernel_only %s  e2e %7.1f  cpu %s  [%s]  max err vs CSR %.1e  halo diff %s\n"
                 % (w, n, v, ms, fr, ("%.4f ms" % ko) if ko else "-", e2e, ("%.1f" % cpu) if cpu else "-", enc, max(errs) if errs else float("nan"),
                    chk.get("exchange_halo_max_abs_diff", "-")))
print(open(os.path.join(ROOT, "profiles", "r02_summary.txt")).read())
