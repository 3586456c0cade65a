#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: tools/show_bench.py file.json [...]"""
import json
import sys
for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline", {})
        print("%s: %s value=%.0f %s  ms/step=%.3f  e2e=%.0f  launches=%s  roofline %.1f %s frac=%.3f" % (
            path, d.get("dtype"), d["value"], d["unit"], d["ms_per_step"], d.get("e2e", {}).get("value", 0), d.get("gpu_launches"),
            r.get("achieved") or 0, r.get("unit"), r.get("frac") or 0))
        for k, v in d.get("kernels", {}).items():
            print("    %-28s %5.1f launches/step  %.3f ms/step" % (k, v["launches_per_step"], v["ms_per_step"]))
        print("    cpu_baseline", d.get("cpu_baseline", {}).get("value"), "clocks", d.get("clocks"))
