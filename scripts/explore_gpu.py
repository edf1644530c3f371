#!/usr/bin/env python
"""Scratch: time the propagate kernel on a reference-built model (needs oracle/_ref harness; not part of bench)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from radiative3d_b200 import workloads as ref_configs  # noqa: E402
from radiative3d_b200 import abi, engine  # noqa: E402
from radiative3d_b200.model import FlatModel  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "halfspace_nearsrc50"
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 9
sizes = [int(float(x)) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1e6", "1e7", "1e8"])]
bits_list = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["-1"])]

with tempfile.TemporaryDirectory() as tmp:
    t0 = time.time()
    env = dict(os.environ, R3D_HARNESS="dump", R3D_HARNESS_OUT=os.path.join(tmp, "m"))
    subprocess.run([os.path.join(ROOT, "oracle/_ref/r3d_ref_harness")] + ref_configs.cmdline(cfg, 10, deg, tmp), cwd=tmp,
                   env=env, check=True, capture_output=True)
    m = FlatModel.load(os.path.join(tmp, "m"))
    print(f"{cfg} deg {deg}: model built+loaded in {time.time() - t0:.1f}s, tables {m.table_bytes() / 1e9:.2f} GB", flush=True)

for bits in bits_list:
    if bits >= 0:
        os.environ["R3D_GUIDE_BITS"] = str(bits)
    t0 = time.time()
    eng = engine.Engine(m)
    print(f"guide_bits={bits}: r3d_create {time.time() - t0:.2f}s", flush=True)
    eng.run_simulation(100000, seed=1)
    eng.sync()
    for n in sizes:
        eng.reset()
        eng.run_simulation(n, seed=2)
        t = eng.sync()
        e, c, k = eng.fetch()
        ev = int(k[abi.R3D_CNT_EVENTS])
        print(f"  n={n:.0e}: {t * 1e3:9.2f} ms  {n / t:.3e} phonons/s  {ev / t:.3e} events/s  ev/ph {ev / n:.2f} "
              f"scat/ph {int(k[5]) / n:.2f} catch/ph {int(k[4]) / n:.4f} lost {int(k[0])} tmo {int(k[1])} inv {int(k[2])}", flush=True)
    eng.close()
