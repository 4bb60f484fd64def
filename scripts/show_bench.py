"""Prints the main keys of a bench.py JSON line (argument: file)."""
import json
import sys

line = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
for k in ("value", "ms_per_step", "roofline", "first_deform", "e2e", "e2e_resident_mesh", "config3",
          "cold_step", "cpu_baseline", "clocks", "gpu_launches"):
    print(k, json.dumps(line.get(k))[:1000])
