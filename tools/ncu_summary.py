"""Summarises an ncu --set full capture (.ncu-rep) into a small text file for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt [workload-name-for-ncu_traffic.json]"""
import csv
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
lines = ["# ncu --set full --clock-control none, file %s" % os.path.basename(rep)]
traffic = None
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append("%-70s %s %s" % (w, r[i], units[i]))
    try:
        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        lines.append("%-70s %.0f byte" % ("dram bytes read+write (traffic per launch)", traffic))
    except Exception as e:  # noqa: BLE001
        lines.append("traffic: n/a (%r)" % (e,))
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if len(sys.argv) > 3 and traffic:
    path = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[sys.argv[3]] = traffic
    json.dump(d, open(path, "w"), indent=1)
