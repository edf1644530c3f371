#!/usr/bin/env python
"""Scratch: one warm-up launch + one measured launch of the propagate kernel, for ncu.
usage: profile_target.py <config> <toa degree> <n phonons>   (model built by integration/_build/r3d_gpu_main)"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from radiative3d_b200 import workloads as ref_configs  # noqa: E402
from radiative3d_b200 import abi, engine  # noqa: E402
from radiative3d_b200.model import FlatModel  # noqa: E402

cfg, deg, n = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3]))
from radiative3d_b200 import reference_host  # noqa: E402
cache = os.path.join(os.environ.get("R3D_MODEL_CACHE", "/tmp/r3d_models"), f"{cfg}_{deg}.r3dmodel")   # one model build per box session
if os.path.exists(cache):
    m = FlatModel.load(cache)
else:
    m = reference_host.build_model(cfg, deg)
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    m.save(cache)
reps = int(os.environ.get("R3D_REPS", "1"))
eng = engine.Engine(m)
eng.run_simulation(n, seed=1)
eng.sync()
best = None
for r in range(reps):
    eng.reset()
    eng.set_profiling(True)
    eng.run_simulation(n, seed=2 + r)
    t = eng.sync()
    best = t if best is None else min(best, t)
if os.environ.get("R3D_TIMING"):
    eng.kernel_times()
e, c, k = eng.fetch()
t = best
print(f"{cfg} deg {deg} n={n}: {t * 1e3:.2f} ms, {n / t:.3e} phonons/s, {int(k[abi.R3D_CNT_EVENTS]) / t:.3e} events/s "
      f"[{int(k[abi.R3D_CNT_EVENTS]) / n:.2f} events, {int(k[abi.R3D_CNT_SCATTERS]) / n:.2f} scatters, {int(k[abi.R3D_CNT_CATCHES]) / n:.3f} catches per phonon]")
