"""The kernel's queue invariants, asserted on the device (libr3dgpu_check.so = the library built with -DR3D_CHECK=1, `make -C
radiative3d_b200/csrc check`): no list outgrows its end of a shared buffer, every slot is in exactly one list after each phase.
A violation traps, the run fails.  Each model runs in its own process, because a process loads one build of the library."""
import os
import subprocess
import sys

import pytest

from conftest import CONFIGS, ROOT

pytestmark = pytest.mark.gpu
CHECK_LIB = os.path.join(ROOT, "radiative3d_b200", "libr3dgpu_check.so")

SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np
from conftest import load_golden
from radiative3d_b200 import abi, engine
m, z = load_golden({cfg!r})
n = {n}
with engine.Engine(m) as eng:
    eng.run_simulation(n, seed=11)
    eng.sync()
    eng.run_simulation(1000, seed=11, first_phonon=n)       # a job smaller than one CTA's slots
    eng.sync()
    e, c, k = eng.fetch()
assert int(k[abi.R3D_CNT_PHONONS]) == n + 1000, k
assert int(k[abi.R3D_CNT_LOST]) + int(k[abi.R3D_CNT_TIMEOUT]) + int(k[abi.R3D_CNT_INVALID]) == n + 1000, k
print("ok", {cfg!r}, int(k[abi.R3D_CNT_EVENTS]))
"""


@pytest.mark.parametrize("cfg", CONFIGS)
def test_queue_invariants(cfg):
    if not os.path.exists(CHECK_LIB):
        pytest.skip("libr3dgpu_check.so not built (make -C radiative3d_b200/csrc check)")
    n = {"halfspace": 3_000_000, "halfspace_nearsrc50": 3_000_000, "crustpinch": 400_000, "lopnor": 300_000, "spherical": 60_000}[cfg]
    env = dict(os.environ, R3D_LIBRARY=CHECK_LIB)
    p = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, cfg=cfg, n=n)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "ok" in p.stdout
