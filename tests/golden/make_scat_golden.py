#!/usr/bin/env python
"""Generate tests/golden/scat_params.npz: the ScatterParams (nu, eps, a, kappa, el, gam0) of every scatterer of the golden
models, in the flattener's order, read from the reference's own objects (oracle/_ref/r3d_ref_harness, mode "scatparams").
Together with golden_<cfg>.npz (the tables the reference built from them) they pin the scatterer-table construction.

    python tests/golden/make_scat_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from radiative3d_b200 import workloads  # noqa: E402
from make_golden import PLAN  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_harness")


def main():
    out = {}
    for cfg, (deg, _) in PLAN.items():
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "sp.txt")
            env = dict(os.environ, R3D_HARNESS="scatparams", R3D_HARNESS_OUT=path)
            p = subprocess.run([HARNESS] + workloads.cmdline(cfg, 10, deg, tmp), cwd=tmp, env=env, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(p.stderr[-2000:])
            out[cfg] = np.loadtxt(path, ndmin=2)
        print(cfg, out[cfg].shape, out[cfg][0])
    np.savez_compressed(os.path.join(HERE, "scat_params.npz"), **out)


if __name__ == "__main__":
    main()
