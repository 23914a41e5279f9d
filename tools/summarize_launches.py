#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name and the share of
libvlmerge's kernels.  usage: summarize_launches.py launches.csv [first_kernel_regex]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), us))
tot = sum(us for _, us in rows)
by = defaultdict(lambda: [0.0, 0])
for name, us in rows:
    by[name][0] += us
    by[name][1] += 1
ours = sum(us for n, us in rows if "vlm::" in n or n.startswith("vlm"))
print(f"{len(rows)} launches, {tot:.0f} us serialised (cold-cache; compare SHARES with bench.py's syrk_share_of_step, not absolutes)")
print(f"libvlmerge kernels' share of the captured launches: {100 * ours / tot:.1f}%")
for name, (us, n) in sorted(by.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{us:10.0f} us {100 * us / tot:5.1f}%  x{n:4d}  {name[:150]}")
