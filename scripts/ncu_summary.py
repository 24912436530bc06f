#!/usr/bin/env python
"""Selected metrics of one ncu report as metric,value,unit CSV:
python scripts/ncu_summary.py rep.ncu-rep > profiles/<name>_ncu_summary.csv"""
import csv, re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
keep = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^lts__t_sector_op_(read|write)_hit_rate\.pct$",
    r"^lts__t_sectors_srcunit_tex_op_read(_lookup_(hit|miss))?\.sum$",
    r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld(_lookup_hit)?\.sum$", r"^l1tex__t_sector_hit_rate\.pct$",
    r"^launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic)$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_(fp64|alu|fma)_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(xu|lsu)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__icc_request_hit_rate\.pct$",
    r"^smsp__average_warp_latency_per_inst_issued\.ratio$",
    r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$",
]
print("metric,value,unit")
for n, u, v in zip(names, units, vals):
    if any(re.search(k, n) for k in keep):
        try:
            if float(v) == 0.0 and "stalled" in n:
                continue
        except ValueError:
            pass
        print(f"{n},{v},{u}")
