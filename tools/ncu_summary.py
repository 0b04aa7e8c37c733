#!/usr/bin/env python
"""Summarise ncu reports (run HERE, no GPU needed): per captured launch the duration, DRAM bytes, tensor-pipe and
memory utilisation -> JSON on stdout.   python tools/ncu_summary.py gpurun_out/prof_tdnn.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "regs",
    "sm__cycles_elapsed.max": "sm_cycles",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
out = []
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    name_i = h.index("Kernel Name")
    for r in rows[2:]:
        d = {"report": rep.split("/")[-1], "kernel": r[name_i][:60]}
        for i, n in enumerate(h):
            if n in WANT:
                v = float(r[i].replace(",", "")) if r[i] else None
                u = units[i]
                if v is not None:
                    if WANT[n].endswith("_MB"):
                        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if WANT[n] == "duration_us":
                        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                d[WANT[n]] = round(v, 3) if v is not None else None
        out.append(d)
print(json.dumps(out, indent=1))
