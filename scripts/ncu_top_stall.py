#!/usr/bin/env python
"""Top SASS instructions by one stall reason, with source lines.  usage: ncu_top_stall.py <rep> <kernel substr> <stall name> [N]"""
import csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym, stall = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; data = rows[2:]
si, ci, ai = hdr.index(stall), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("R3D_PROFILE_LIB", os.path.join(ROOT, "radiative3d_b200", "libr3dgpu.so"))], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cands, lines, cur, grab = [], [], None, False
for l in sass.splitlines():
    if l.startswith("//--------------------- .text."):
        if grab and lines: cands.append(lines)
        grab = sym in l and "_ZN" in l; lines, cur = [], None; continue
    if not grab: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): lines.append(cur)
if grab and lines: cands.append(lines)
lines = min(cands, key=lambda c: abs(len(c) - len(data))) if cands else []
if len(lines) != len(data): print(f"warning: cubin {len(lines)} vs report {len(data)} instructions", file=sys.stderr); lines = (lines + [None] * len(data))[:len(data)]
tot = sum(int(r[si] or 0) for r in data)
order = sorted(range(len(data)), key=lambda k: -int(data[k][si] or 0))[:top]
for k in order:
    r = data[k]; w = lines[k]
    print(f"{k:6d} {100 * int(r[si] or 0) / max(tot, 1):6.2f}% exec {int(r[ci]):10d} lanes {r[ai]:>5s}  {(w[0] + ':' + str(w[1])) if w else '?':26s} {r[1].strip()[:60]}")
