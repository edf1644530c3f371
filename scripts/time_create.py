#!/usr/bin/env python
"""Scratch: wall-clock of r3d_create / run / fetch / destroy on the bench workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiative3d_b200 import engine, reference_host
m = reference_host.build_model("halfspace_nearsrc50", 9)
for i in range(3):
    t0 = time.perf_counter(); e = engine.Engine(m); t1 = time.perf_counter()
    e.run_simulation(10_000_000, seed=i); e.sync(); t2 = time.perf_counter()
    e.fetch(); t3 = time.perf_counter(); e.close(); t4 = time.perf_counter()
    print(f"create {t1-t0:.3f}s run {t2-t1:.3f}s fetch {t3-t2:.3f}s destroy {t4-t3:.3f}s", flush=True)
