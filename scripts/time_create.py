#!/usr/bin/env python
"""Scratch: wall-clock of r3d_create / run / fetch / destroy on the bench workload, host arrays pinned as in bench.py
(R3D_TIMING=1 prints the laps inside r3d_create)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from radiative3d_b200 import engine, reference_host
from radiative3d_b200.model import _ARRAYS
m = reference_host.build_model("halfspace_nearsrc50", 9)
keep = []
for name, _ in _ARRAYS:
    t = torch.from_numpy(getattr(m, name)).pin_memory(); keep.append(t); setattr(m, name, t.numpy())
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
for i in range(4):
    t0 = time.perf_counter(); e = engine.Engine(m); t1 = time.perf_counter()
    e.run_simulation(n, seed=i); dev = e.sync(); t2 = time.perf_counter()
    e.fetch(); t3 = time.perf_counter(); e.close(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  run+sync {1e3*(t2-t1):.2f} ms (device {1e3*dev:.2f})  fetch {1e3*(t3-t2):.2f} ms  destroy {1e3*(t4-t3):.2f} ms  total {1e3*(t4-t0):.2f}", flush=True)
