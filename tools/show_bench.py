#!/usr/bin/env python
"""Prints the key fields of bench.py JSON lines read from stdin or files (development helper)."""
import json
import sys

texts = [open(p).read() for p in sys.argv[1:]] or [sys.stdin.read()]
for text in texts:
    for line in text.strip().splitlines():
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("impl") == "reference":
            print("reference:", d["value"], d["unit"], d["cpu_baseline"]["cores"], "cores")
            continue
        r = d["roofline"]
        print(f"N={d['n_gpus']} value={d['value']} e2e={d['e2e']['value']} ms/step={d['ms_per_step']} launches={d['gpu_launches']}")
        print(f"  syrk: {r['achieved']} TF/s frac={r['frac']} share={r['syrk_share_of_step']} traffic={r.get('traffic')}")
        for k, v in r["by_shape"].items():
            print(f"    {k}: {v}")
        m = d.get("merge") or {}
        if m:
            print(f"  merge: {m['value']} GB/s frac={m['roofline']['frac']} e2e={m['e2e']}")
        for k in ("reference_hook_on_gpu", "regmean", "gram_file", "fused_route", "irtr", "vitl", "forward_variants", "gram_parity_rel_fro_reduced", "cpu_baseline", "gram_parity_rel_fro", "clocks"):
            if d.get(k):
                print(f"  {k}: {d[k]}")
