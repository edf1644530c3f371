#!/usr/bin/env python
"""Scratch: warp-time by kind of chunk (R3D_TIMING=1 makes r3d_kernel_times print it).  usage: chunk_clocks.py <config> <deg> <n>"""
import os, sys
os.environ["R3D_TIMING"] = "1"
_clk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "radiative3d_b200", "libr3dgpu_clocks.so")   # make -C radiative3d_b200/csrc clocks
if os.path.exists(_clk):
    os.environ.setdefault("R3D_LIBRARY", _clk)
else:
    print("chunk_clocks.py: libr3dgpu_clocks.so not built (make -C radiative3d_b200/csrc clocks): per-chunk figures will be zero", file=sys.stderr)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiative3d_b200 import abi, engine, reference_host
cfg, deg, n = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3]))
m = reference_host.build_model(cfg, deg)
eng = engine.Engine(m)
eng.run_simulation(n, seed=1); eng.sync(); eng.reset(); eng.set_profiling(True)
eng.run_simulation(n, seed=2)
t = eng.sync()
kt = eng.kernel_times()
print(f"{cfg} deg {deg} n={n}: {t * 1e3:.2f} ms, {n / t:.3e} phonons/s; {kt}")
