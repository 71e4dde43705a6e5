#!/usr/bin/env python
"""Stall samples of an .ncu-rep aggregated per CUDA source line (needs -lineinfo and --import-source on).
usage: python tools/srclines.py file.ncu-rep [top]"""
import collections, csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if "# Samples" in r)
data = rows[rows.index(hdr) + 1:]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
samples, execd, text = collections.Counter(), collections.Counter(), {}
cur_file = ""
for r in data:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ix["# Samples"]] or 0); e = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    key = r[ix["Line No"]]
    samples[key] += s; execd[key] += e; text.setdefault(key, r[ix["Source"]].strip())
tot = sum(samples.values()); tote = sum(execd.values())
print("total samples", tot, "instructions executed", tote)
for key, s in samples.most_common(top):
    print(f"{100*s/tot:6.2f}% smp {100*execd[key]/max(tote,1):6.2f}% ins  L{key:>5s}  {text[key][:110]}")
