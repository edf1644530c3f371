#!/usr/bin/env python
"""Generate tests/golden/events_<cfg>.npz: the reference's own event reports (--reports=ALL_ON, dataout.cpp:484-617) for runs
of its GenerateEventPhonon()+Propagate() loop on the Philox draw stream (oracle/_ref/r3d_ref_harness, mode "run"), i.e. for the
very phonons the GPU traces with the same seed.

    python tests/golden/make_event_golden.py
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from radiative3d_b200 import workloads  # noqa: E402
from make_golden import PLAN as MODEL_PLAN, SEED  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_harness")
KINDS = ("GEN", "SCT", "COL", "REF", "CEL", "LST", "TMO", "INV")
PLAN = {"halfspace": 80, "halfspace_nearsrc50": 40, "crustpinch": 25, "lopnor": 25, "spherical": 6}

LINE = re.compile(r"^(\w{3}):\s+(\d+)\s+([PS])\s+ttpl:\(\s*(\S+)\s+(\S+)\s*\)\s+xyz:\(\s*(\S+)\s+(\S+)\s+(\S+)\s*\)\s+thph:\(\s*(\S+)\s+(\S+)\s*\)"
                  r"\s+a:\(\s*(\S+)\s*\)\s+cell:\s*(\S+)\s+it:\s*(\d+)")

EVENT_LINE_DTYPE = np.dtype([("kind", "u1"), ("type", "u1"), ("sid", "<u8"), ("time", "<f8"), ("pathlen", "<f8"), ("xyz", "<f8", (3,)),
                             ("theta", "<f8"), ("phi", "<f8"), ("amp", "<f8"), ("it", "<u4"), ("phonon", "<u4")])


def parse_reports(path):
    """reports.dat -> structured array, one row per line; `phonon` counts GEN lines (0, 1, ...)."""
    rows, ph = [], -1
    for line in open(path):
        m = LINE.match(line)
        if not m:
            continue
        k = KINDS.index(m.group(1))
        if k == 0:
            ph += 1
        v = [float(m.group(i)) for i in range(4, 12)]
        rows.append((k, 0 if m.group(3) == "P" else 1, int(m.group(2)), v[0], v[1], (v[2], v[3], v[4]), v[5], v[6], v[7], int(m.group(13)), max(ph, 0)))
    return np.array(rows, dtype=EVENT_LINE_DTYPE)


def report_args(cfg, n, deg, outdir):
    return [a for a in workloads.cmdline(cfg, n, deg, outdir) if not a.startswith("--reports")] + ["--reports=ALL_ON"]


def main():
    if not os.path.exists(HARNESS):
        sys.exit("oracle/_ref/r3d_ref_harness is missing: run `make -C oracle ref` where /root/reference exists")
    for cfg, n in PLAN.items():
        deg = MODEL_PLAN[cfg][0]
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, R3D_HARNESS="run", R3D_HARNESS_OUT=os.path.join(tmp, "r"), R3D_HARNESS_SEED=str(SEED))
            p = subprocess.run([HARNESS] + report_args(cfg, n, deg, tmp), cwd=tmp, env=env, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(p.stderr[-2000:])
            ev = parse_reports(os.path.join(tmp, "reports.dat"))
        assert ev["phonon"].max() == n - 1
        path = os.path.join(HERE, f"events_{cfg}.npz")
        np.savez_compressed(path, events=ev, n=np.int64(n), seed=np.uint64(SEED), toa_degree=np.int64(deg))
        counts = {KINDS[k]: int((ev["kind"] == k).sum()) for k in range(8)}
        print(f"events_{cfg}.npz: {n} phonons, {ev.size} lines {counts}, {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
