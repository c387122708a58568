"""Aggregates the source page of an ncu report by source line: warp instructions executed, stall samples.
Usage: python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
recs = []
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        recs.append((cur_file, int(r[0]), r[1].strip(), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                     int(d.get("Thread Instructions Executed") or 0)))
tot = sum(x[3] for x in recs) or 1
tots = sum(x[4] for x in recs) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
for f, ln, src, n, smp, tn in sorted(recs, key=lambda x: -x[3])[:top]:
    print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  %s:%d  %s" % (100.0 * n / tot, 100.0 * smp / tots, tn / max(n, 1), f, ln, src[:110]))
if len(sys.argv) > 3:   # per-line dump of one file in line order: python tools/ncu_lines.py rep N file
    for f, ln, src, n, smp, tn in sorted(recs, key=lambda x: (x[0], x[1])):
        if f == sys.argv[3] and n:
            print("%6.2f%% %5.1f%% smp lanes %4.1f  %d  %s" % (100.0 * n / tot, 100.0 * smp / tots, tn / max(n, 1), ln, src[:120]))
