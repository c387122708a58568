"""Per-source-line share of executed instructions and stall samples from an ncu report captured with
--import-source on (reads `ncu -i REP --page source --print-source cuda,sass --csv`).
Usage: python tools/ncu_lines.py REPORT.ncu-rep [min_share_pct]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
inst, samp, text = collections.Counter(), collections.Counter(), {}
for r in rows:
    if "Instructions Executed" in r:
        if hdr is not None:
            break          # first kernel only
        hdr = r
        ci, si = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if hdr is None or len(r) <= ci or not r[0].isdigit():
        continue
    ln = int(r[0])
    text[ln] = r[1]
    if r[ci].isdigit():
        inst[ln] += int(r[ci])
        samp[ln] += int(r[si]) if r[si].isdigit() else 0
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
for ln in sorted(inst):
    if 100.0 * inst[ln] / ti >= thr or 100.0 * samp[ln] / ts >= thr:
        print("%5d  inst %5.1f%%  samples %5.1f%%  %s" % (ln, 100.0 * inst[ln] / ti, 100.0 * samp[ln] / ts, text[ln].strip()[:120]))
