#!/usr/bin/env python
"""The drop-in program end to end: integration/_build/r3d_gpu_main on a BASELINE workload at TOA degree
9, whole-process wall clock and its own RunSimulation breakdown (flatten / r3d_create from pageable memory / loop / fetch).
usage: dropin_e2e.py <config> <n phonons> [devices, e.g. 0,1]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiative3d_b200 import reference_host
cfg, n = sys.argv[1], int(float(sys.argv[2]))
devices = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (0,)
with tempfile.TemporaryDirectory() as tmp:
    t = time.perf_counter()
    p = reference_host.run(cfg, n, 9, tmp, seed=1, devices=devices)
    wall = time.perf_counter() - t
    if p.returncode != 0:
        sys.exit(p.stderr[-2000:])
    lines = [ln for ln in p.stderr.splitlines() if ln.startswith("r3d-gpu:")]
    nfiles = len([f for f in os.listdir(tmp) if f.startswith("seis_")])
print(f"{cfg} n={n} devices={devices}: whole process {wall:.2f} s wall (model build by the reference's host code, scatterer tables on the GPU, "
      f"GPU loop, {nfiles} output files)")
for ln in lines:
    print("   ", ln)
