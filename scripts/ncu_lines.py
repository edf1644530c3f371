#!/usr/bin/env python
"""Attribute an .ncu-rep's per-SASS-instruction counts to CUDA source lines (run here, no GPU needed).

    python scripts/ncu_lines.py <file.ncu-rep> <kernel name substring in the cubin symbol> [top N]

ncu's CSV source page is SASS-only, so the i-th instruction of the profiled kernel is matched with the i-th
instruction of `nvdisasm --print-line-info` on the cubin inside radiative3d_b200/libr3dgpu.so (built with -lineinfo).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
inst = [(r[1].strip(), int(r[ci]), int(r[ct]), int(r[cs])) for r in rows[2:] if len(r) > ct]

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("R3D_PROFILE_LIB", os.path.join(ROOT, "radiative3d_b200", "libr3dgpu.so"))], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

# every function whose symbol contains `sym`; keep the one whose instruction count equals the report's
cands, lines, cur, grab = [], [], None, False
for l in sass.splitlines():
    if l.startswith("//--------------------- .text."):
        if grab and lines:
            cands.append(lines)
        grab = sym in l and "_ZN" in l
        lines, cur = [], None
        continue
    if not grab:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
if grab and lines:
    cands.append(lines)
lines = min(cands, key=lambda c: abs(len(c) - len(inst))) if cands else []
if len(lines) != len(inst):
    print(f"warning: {len(lines)} instructions in the cubin vs {len(inst)} in the report; attribution may be shifted", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (txt, ni, nt, ns), where in zip(inst, lines):
    a = agg[where]
    a[0] += ni; a[1] += nt; a[2] += ns; a[3] += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp instructions {tot_i}, stall samples {tot_s}, SASS instructions {len(inst)}")
src_cache = {}
def text(where):
    if not where: return ""
    f, n = where
    for d in ("radiative3d_b200/csrc", "include"):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:90] if n - 1 < len(src_cache[p]) else ""
    return ""
print(f"{'file:line':28s} {'warp-inst%':>10s} {'lanes':>6s} {'stall%':>7s}  source")
for where, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    name = f"{where[0]}:{where[1]}" if where else "?"
    print(f"{name:28s} {100 * a[0] / tot_i:10.2f} {a[1] / max(a[0], 1):6.1f} {100 * a[2] / max(tot_s, 1):7.2f}  {text(where)}")

# optional: dump the SASS of one source line:  ... <top N> <file:line>
if len(sys.argv) > 4:
    f, n = sys.argv[4].split(":")
    for (txt, ni, nt, ns), where in zip(inst, lines):
        if where == (f, int(n)):
            print(f"{ni:10d} {nt / max(ni, 1):5.1f} {ns:5d}  {txt}")
