#!/usr/bin/env python
"""profiles/ncu_by_line.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTR [top]

Joins `ncu --page source --csv` (per SASS address: instructions executed, stall samples) with
`nvdisasm -g` line info of the same object file and prints the totals per source line, so a
kernel's instruction budget can be read against the CUDA source.  Analysis tooling only."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line = {}
inside = False
cur = None
for line in dis.splitlines():
    if line.startswith("//---") and ".text." in line:
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = {n: i for i, n in enumerate(rows[hdr])}
base = None
per_line = defaultdict(lambda: [0, 0, 0])
per_op = defaultdict(int)
total = 0
for r in rows[hdr + 1:]:
    if len(r) <= cols["Instructions Executed"]:
        continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if base is None:
        base = a
    ex = int(float(r[cols["Instructions Executed"]] or 0))
    smp = int(float(r[cols["# Samples"]] or 0))
    key, op = addr2line.get(a - base, (("?", 0), "?"))
    per_line[key][0] += ex
    per_line[key][1] += smp
    per_line[key][2] += 1
    per_op[op.split(".")[0]] += ex
    total += ex
print("total warp instructions executed: %d" % total)
srcs = {}
for (f, l), (ex, smp, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
    print("%5.1f%% %11d smp %6d  %s:%d  %s" % (100.0 * ex / total, ex, smp, f, l, text[:90]))
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / total) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1])[:16]))
