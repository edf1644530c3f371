#!/usr/bin/env python
"""Scratch: which fields of which phonons differ between the GPU trace and an INV fixture."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, INV_CASES
from radiative3d_b200 import engine
for case in (sys.argv[1:] or INV_CASES):
    m, z = load_golden(case, prefix="inv")
    n, seed = int(z["run_n"]), int(z["run_seed"])
    ref = z["run_finals"]
    with engine.Engine(m) as eng:
        fin = eng.trace(n, seed)
        e, c, k = eng.fetch()
    print(case, "counters gpu", k, "ref", z["run_counters"])
    bad = np.zeros(n, dtype=bool)
    for f in ("moves", "cell", "type", "fate", "draws"):
        d = fin[f] != ref[f]
        bad |= d
        print("  ", f, int(d.sum()))
    for i in np.nonzero(bad)[0][:6]:
        print("   gpu", i, fin[i]); print("   ref", i, ref[i])
