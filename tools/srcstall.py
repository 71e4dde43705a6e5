#!/usr/bin/env python
"""Per-instruction stall samples of an .ncu-rep (source page): top instructions by samples, and totals by region.
usage: python tools/srcstall.py file.ncu-rep [top]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print({k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100*s/tot:5.2f}% exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:70]:70s} {st}")
# region totals: contiguous ranges given as extra args lo:hi
for a in sys.argv[3:]:
    lo, hi = map(int, a.split(":"))
    s = sum(int(r[ix["# Samples"]] or 0) for r in data[lo:hi + 1])
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in data[lo:hi + 1])
    print(f"region {lo}:{hi} samples {100*s/tot:.1f}% executed {ex}")
