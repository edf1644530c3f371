#!/usr/bin/env python
"""Stall-reason samples of an .ncu-rep by SASS block.  usage: ncu_stalls.py <rep> [block size]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; blk = int(sys.argv[2]) if len(sys.argv) > 2 else 500
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in names}
ci = hdr.index("Instructions Executed")
data = rows[2:]
tot = {h: 0 for h in names}
print("block      inst%  " + " ".join(f"{h[6:12]:>7s}" for h in names))
ti = sum(int(r[ci]) for r in data)
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    v = {h: sum(int(r[idx[h]] or 0) for r in seg) for h in names}
    for h in names: tot[h] += v[h]
    print(f"{b:6d} {100 * sum(int(r[ci]) for r in seg) / ti:8.2f}  " + " ".join(f"{v[h]:7d}" for h in names))
print("total            " + " ".join(f"{tot[h]:7d}" for h in names))
