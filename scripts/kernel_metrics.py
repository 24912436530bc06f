#!/usr/bin/env python
"""Fill profiles/kernel_metrics.json (read by bench.py's roofline block) from an ncu report:

    python scripts/kernel_metrics.py gpurun_out/prof.ncu-rep <stars per launch> f64_f64 <committed summary name>
"""
import csv
import json
import os
import subprocess
import sys

rep, stars, key, source = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
val = dict(zip(rows[0], rows[2]))
unit = dict(zip(rows[0], rows[1]))


def num(name):
    return float(val[name].replace(",", ""))


def to_bytes(name):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[name]]
    return num(name) * scale


m = {
    "dram_bytes_per_star": (to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")) / stars,
    "issue_active_frac": num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
    "fp64_pipe_frac": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100,
    "lsu_data_pipe_frac": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100,
    "instructions_per_star": 32 * num("smsp__inst_executed.sum") / stars,
    "l2_read_hit_rate": num("lts__t_sector_op_read_hit_rate.pct") / 100,
    "registers_per_thread": int(num("launch__registers_per_thread")),
    "stars_per_launch": stars,
    "source": source,
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                    "kernel_metrics.json")
try:
    allm = json.load(open(path))
except Exception:
    allm = {}
allm[key] = m
json.dump(allm, open(path, "w"), indent=1)
print(json.dumps(m, indent=1))
