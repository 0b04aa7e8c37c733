#!/usr/bin/env python
"""Readable summary of bench.py JSON lines found in the given files (default: gpurun tail in /tmp/gpu_out.txt)."""
import json
import sys

for path in (sys.argv[1:] or ["/tmp/gpu_out.txt"]):
    for l in open(path):
        if l.startswith("{"):
            d = json.loads(l)
            if "roofline" not in d:
                print(l.strip()[:300]); continue
            print("value %.1fM  ms/step %.4f  e2e %.1fM (%.3f ms)  clocks %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["clocks"]))
            r = d["roofline"]
            print(" tdnn agg: %.1f TF frac %.3f share %.3f stepfrac %.3f" % (r["achieved"], r["frac"], r["share_of_step"], r["step_frac_of_tensor_peak"]))
            for k in r["launches"]:
                print("   %-28s %.4f ms  %8.1f %s  frac %.3f" % (k["kernel"], k["ms"], k["achieved"], k["unit"], k["frac"]))
            if "cpu_baseline" in d:
                print(" cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"])
        elif any(w in l for w in ("passed", "failed", "gpurun]", "rror")):
            print(l.strip())
