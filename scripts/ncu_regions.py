#!/usr/bin/env python
"""Aggregate executed warp instructions of a kernel by named source-line ranges.
usage: ncu_regions.py rep kernel-substring file:lo-hi=name ..."""
import csv, subprocess, sys
from collections import defaultdict
rep, want = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    loc, name = a.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    regions.append((f, int(lo), int(hi), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fn = file = hdr = None
agg = defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for r in rows:
    if not r: continue
    if r[0] == "File Path": file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or r[0] in ("", "-") or want not in fn: continue
    try: line = int(r[0]); inst = int(r[ii]); ti = int(r[it]); sm = int(r[isamp])
    except ValueError: continue
    name = file
    for f, lo, hi, n in regions:
        if f == file and lo <= line <= hi: name = n; break
    a = agg[name]; a[0] += inst; a[1] += ti; a[2] += sm
    tot[0] += inst; tot[1] += ti; tot[2] += sm
print(f"total warp inst {tot[0]}  thread inst {tot[1]}  samples {tot[2]}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*v[0]/tot[0]:5.1f}% inst  {100*v[2]/max(tot[2],1):5.1f}% smp  thr/inst {v[1]/max(v[0],1):4.1f}  {k}")
