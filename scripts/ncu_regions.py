#!/usr/bin/env python
"""Executed warp instructions of an .ncu-rep by consecutive SASS blocks (to see which part of a fused kernel is hot).
usage: ncu_regions.py <rep> <kernel substring> [block size]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym = sys.argv[1], sys.argv[2]
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 250
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
inst = [(r[1].strip(), int(r[ci]), int(r[ct]), int(r[cs])) for r in rows[2:] if len(r) > ct]
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("R3D_PROFILE_LIB", os.path.join(ROOT, "radiative3d_b200", "libr3dgpu.so"))], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cands, lines, cur, grab = [], [], None, False
for l in sass.splitlines():
    if l.startswith("//--------------------- .text."):
        if grab and lines: cands.append(lines)
        grab = sym in l and "_ZN" in l
        lines, cur = [], None
        continue
    if not grab: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): lines.append(cur)
if grab and lines: cands.append(lines)
lines = min(cands, key=lambda c: abs(len(c) - len(inst)))
tot = sum(i[1] for i in inst); tots = sum(i[3] for i in inst)
print(f"total {tot} warp inst, {len(inst)} SASS")
for b in range(0, len(inst), blk):
    seg = inst[b:b + blk]
    ni = sum(i[1] for i in seg); nt = sum(i[2] for i in seg); ns = sum(i[3] for i in seg)
    where = collections.Counter(w for w in lines[b:b + blk] if w and "resident" in w[0])
    top = ", ".join(f"{w[1]}" for w, _ in where.most_common(4))
    print(f"{b:6d} {100 * ni / tot:6.2f}% lanes {nt / max(ni, 1):5.1f} stall {100 * ns / max(tots, 1):5.2f}%  resident lines: {top}")
if len(sys.argv) > 5:   # dump a SASS index range: <rep> <sym> <blk> <from> <to>
    a, b = int(sys.argv[4]), int(sys.argv[5])
    last = None
    for k in range(a, min(b, len(inst))):
        txt, ni, nt, ns = inst[k]
        w = lines[k]
        tag = f"{w[0]}:{w[1]}" if w else "?"
        print(f"{k:6d} {ni:10d} {nt / max(ni, 1):5.1f} {ns:6d}  {tag if tag != last else '':28s} {txt[:70]}")
        last = tag
