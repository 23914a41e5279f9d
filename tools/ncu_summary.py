#!/usr/bin/env python
"""ncu-rep -> summary CSV (metric,unit,value per line) for profiles/: keeps the metrics the roofline discussion uses.
usage: ncu_summary.py report.ncu-rep [launch_index] > profiles/<name>.summary.csv"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
keep = re.compile(r"^(Kernel Name|Block Size|Grid Size|dram__|gpu__time_duration|gpu__dram_throughput|sm__pipe_tensor|"
                  r"sm__inst_executed_pipe_tensor|sm__throughput|sm__warps_active|launch__|lts__t_sector_hit_rate|"
                  r"lts__t_bytes\.sum|l1tex__m_xbar2l1tex_read_bytes\.sum|smsp__cycles_active\.avg|sm__cycles_elapsed\.(avg|max)$)")
for h, u, v in zip(hdr, units, vals):
    if keep.match(h):
        print(f"{h},{u},{v}")
