#!/usr/bin/env python
"""profiles/traffic.json from ncu captures: DRAM bytes per launch of the FIR kernel.
usage: scripts/ncu_traffic.py WL:kernel=path.ncu-rep[:streams] ...   (e.g. C3:tensor=gpurun_out/prof_C3.ncu-rep)
Each report is an `ncu --set full --clock-control none` capture of ONE launch; the numbers are
dram__bytes_read.sum / dram__bytes_write.sum of that launch. bench.py reads the file for
roofline.traffic (it never profiles itself)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "traffic.json")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
data = json.load(open(OUT)) if os.path.exists(OUT) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    path, _, streams = rest.partition(":")
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    col = {h: i for i, h in enumerate(rows[0])}

    def metric(name):
        return float(rows[2][col[name]]) * UNIT[rows[1][col[name]]]
    data[key] = {"dram_read_bytes": metric("dram__bytes_read.sum"), "dram_write_bytes": metric("dram__bytes_write.sum"),
                 "kernel_name": rows[2][col["Kernel Name"]][:80], "duration_us_under_ncu": float(rows[2][col["gpu__time_duration.sum"]]),
                 "source": f"{os.path.relpath(path, ROOT)} (ncu --set full --clock-control none, one launch) via scripts/ncu_traffic.py"}
    if streams:
        data[key]["streams"] = int(streams)
json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
print(json.dumps(data, indent=1, sort_keys=True))
