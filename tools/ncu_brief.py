"""Prints the handful of ncu raw metrics that matter for the HBM-bound kernels of this repo.
Usage: python tools/ncu_brief.py report.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.sum ", "l1tex__data_pipe_lsu_wavefronts.sum.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_lgds", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum ", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled"]
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2 + idx]
print(data[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for h, u, v in zip(hdr, units, data):
    if any(k.strip() in h for k in KEYS):
        print("%-90s %-14s %s" % (h, u, v))
