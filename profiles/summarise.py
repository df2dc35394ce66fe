#!/usr/bin/env python
"""profiles/summarise.py -- turn gpurun_out/launches.csv (+ *.ncu-rep) into the committed summaries.

usage: python profiles/summarise.py <round tag>      e.g. r01
"""
import csv
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*lambda.*", "<lambda>", name)
    return name


launches = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if not os.path.exists(launches):
    launches = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[4]) + " " + r[7] + "x" + r[8]
        ns = float(r[14])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    out.append("## Launch list (ncu --metrics gpu__time_duration.sum --clock-control none), %d launches, %.2f ms total\n" % (len(rows), total / 1e6))
    out.append("bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary: the bench workload (batch 4096): device-API round-trip steps, one profiled pass, host-API encode and decode steps in sub-chunks, synth. Cold-cache, serialised: compare shares.\n")
    # the device-API encode steps alone (grids over 4096 images / 8192 planes): the shares bench.py's kernel table must agree with
    dev = collections.OrderedDict()
    for r in rows:
        g = [int(x) for x in re.findall(r"\d+", r[8])]
        nm = short(r[4])
        if not (4096 in g or 8192 in g or 1024 in g) or nm.startswith("void at::") or nm in ("k_synth",):
            continue
        a = dev.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14])
    dtot = sum(v[1] for v in dev.values())
    if dtot:
        out.append("### Device-API round-trip steps only (grids over 4096 images / 8192 planes / 1024 four-stream blocks), share of the step\n")
        out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, (n, ns) in sorted(dev.items(), key=lambda kv: -kv[1][1])[:25]:
            out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / dtot))
        out.append("")
    out.append("| kernel block x grid | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / total))
    out.append("")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_*.ncu-rep" % tag))) or sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep")))
for rep in reps:
    try:
        txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL, text=True)
    except Exception as e:
        out.append("## %s: could not be read (%s)\n" % (os.path.basename(rep), e))
        continue
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        out.append("## ncu --set full: %s -- `%s`\n" % (os.path.basename(rep), short(vals[hdr.index("Kernel Name")])))
        out.append("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append("| %s | %s | %s |" % (w, vals[i], units[i]))
        out.append("")
open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag), "w").write("# ncu summaries, round tag %s\n\n" % tag + "\n".join(out) + "\n")
print("\n".join(out)[:6000])
