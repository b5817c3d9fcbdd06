#!/usr/bin/env python3
"""Extracts the judged numbers from an ncu report: for every kernel instance a row of selected raw-page metrics (CSV on stdout).
Usage: ncu_extract.py report.ncu-rep > profiles/rNN_ncu_raw_selected.csv"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tma.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True).stdout.decode("utf-8", "replace")
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
cols = [hdr.index(w) for w in WANT if w in hdr]
out = csv.writer(sys.stdout)
out.writerow(["Kernel Name"] + [hdr[c] for c in cols])
out.writerow([""] + [units[c] for c in cols])
for r in rows[2:]:
    if len(r) > ik:
        out.writerow([r[ik][:60]] + [r[c] for c in cols])
