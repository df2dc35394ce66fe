#!/usr/bin/env python
"""profiles/ncu_traffic.py [round tag] -- DRAM traffic of the kernels captured with `ncu --set full` (profiles/run_ncu_r02.sh:
gpurun_out/prof_<tag>_<kernel>.ncu-rep, one launch each over `--batch 1024`) -> profiles/ncu_traffic.json, which bench.py reads
for `roofline.traffic`.  dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, divided by the images it covered."""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
BATCH = 1024
LABEL = {"k_front_luma": "k_front_luma", "kd_backend": "d_backend", "kd_inv_rows_t": "d_inv_rows_t", "kd_serial_front": "d_serial_front",
         "k_e16_residual": "y_e16_residual", "k_ll2_code": "y_ll2_code", "k_entropy": "entropy_pack", "k_e20_bands": "y_e20_cleanup",
         "kd_y_markers": "d_markers_y", "k_idwt_level": "k_idwt_level<256>", "k_dwt_level": "k_dwt_level<256>"}
TIME_US = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_*.ncu-rep" % tag))):
    kern = re.sub(r"^prof_%s_|\.ncu-rep$" % tag, "", os.path.basename(rep))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr is None or len(rows) < hdr + 3:
        continue
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    col = {n: i for i, n in enumerate(names)}

    def get(m):
        i = col[m]
        return float(vals[i].replace(",", "")) * UNIT.get(units[i], 1)
    total = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    out[LABEL.get(kern, kern)] = {"dram_bytes_per_image": round(total / BATCH, 1), "dram_read_bytes": get("dram__bytes_read.sum"),
                                  "dram_write_bytes": get("dram__bytes_write.sum"), "launch_us": round(float(vals[col["gpu__time_duration.sum"]].replace(",", "")) *
                                                   TIME_US.get(units[col["gpu__time_duration.sum"]], 1.0), 2),
                                  "images_in_launch": BATCH, "source": "profiles/%s_ncu_summary.md (%s)" % (tag, os.path.basename(rep))}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
