#!/usr/bin/env python
"""Rank CUDA source lines of an .ncu-rep by executed instructions / stall samples.

usage: ncu_lines.py report.ncu-rep [kernel-substring] [top]
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = cur_fn = None
hdr = None
agg = defaultdict(lambda: [0, 0, 0, ""])  # inst, thread inst, samples, text
tot = defaultdict(lambda: [0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_tinst = hdr.index("Thread Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("", "-") or want not in (cur_fn or ""):
        continue
    try:
        line = int(r[0])
        inst = int(r[i_inst])
        tinst = int(r[i_tinst])
        samp = int(r[i_samp])
    except ValueError:
        continue
    key = (cur_fn, cur_file, line)
    a = agg[key]
    a[0] += inst
    a[1] += tinst
    a[2] += samp
    a[3] = r[1].strip()[:110]
    tot[cur_fn][0] += inst
    tot[cur_fn][1] += samp
for fn, (ti, ts) in tot.items():
    print(f"== {fn}: {ti} warp instructions, {ts} samples")
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    items.sort(key=lambda kv: -kv[1][2 if "--by-samples" in sys.argv else 0])
    for (f, file, line), (inst, tinst, samp, text) in items[:top]:
        print(f"{100*inst/ti:5.1f}% inst {100*samp/max(ts,1):5.1f}% smp  thr/inst {tinst/max(inst,1):4.1f}  {file}:{line}  {text}")
