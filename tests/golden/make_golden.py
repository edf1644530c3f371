#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/r3d_ref_harness).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):

    python tests/golden/make_golden.py

For each BASELINE.json workload (reduced take-off-angle degree and phonon count so the fixtures
stay small) this stores
  * the flattened model the reference built (cells, faces, scatterer / source CDFs, seismometers),
  * deterministic sub-kernel vectors made by calling the reference's own methods
    (GetPathToBoundary, AdvanceLength, ProbDist::GetRandomIndex, Seismometer::CatchPhonon),
  * a whole run of the reference's GenerateEventPhonon()+Propagate() loop with rand() replaced by
    the Philox4x32-10 stream (seed, phonon index, draw ordinal): per-phonon end states, bins, counters.
and, model-independent, golden_free.npz: Phonon::Transform, RTCoef on random interfaces, and the
reference's built-in --rtcoef-test table (rtcoef.cpp:687-742).
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

from radiative3d_b200 import workloads as ref_configs  # noqa: E402
from radiative3d_b200 import abi  # noqa: E402
from radiative3d_b200.model import FlatModel, _ARRAYS  # noqa: E402
from oracle_binding import load_bins  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_harness")
SEED = 20261018

# config -> (toa degree, phonons in the whole-run fixture)
PLAN = {
    "halfspace": (3, 4000),
    "halfspace_nearsrc50": (2, 2000),
    "crustpinch": (2, 1500),
    "lopnor": (2, 1500),
    "spherical": (2, 300),
}


def harness(mode, args, out, cwd, extra_env=None):
    env = dict(os.environ, R3D_HARNESS=mode, R3D_HARNESS_OUT=out, R3D_HARNESS_SEED=str(SEED))
    env.update(extra_env or {})
    p = subprocess.run([HARNESS] + args, cwd=cwd, env=env, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"harness {mode} failed:\n{p.stderr[-2000:]}")
    return p


def pairs(path, win, wout):
    """The harness writes one input line then one output line per case."""
    rows = [np.array(l.split(), dtype=np.float64) for l in open(path) if l.strip()]
    x = np.array(rows[0::2]).reshape(-1, win)
    y = np.array(rows[1::2]).reshape(-1, wout)
    assert x.shape[0] == y.shape[0] and x.shape[0] > 0
    return x, y


def model_fields(m):
    d = {name: getattr(m, name) for name, _ in _ARRAYS}
    d["scalars_f"] = np.array([m.freq_hz, m.ttl, m.bin_dt, *m.earth_center, m.min_theta, m.max_theta, m.slow_concern,
                               *m.src_loc, m.cyl_radius2])
    d["scalars_i"] = np.array([m.loop_concern, m.n_bins, m.ecs_radial, m.no_deflect, m.src_cell, m.cell_kind], dtype=np.int64)
    return d


def main():
    if not os.path.exists(HARNESS):
        sys.exit("oracle/_ref/r3d_ref_harness is missing: run `make -C oracle ref` where /root/reference exists")
    with tempfile.TemporaryDirectory() as tmp:
        harness("vectors-free", [], tmp, tmp)
        tx, ty = pairs(os.path.join(tmp, "transform.txt"), 6, 3)
        rx, ry = pairs(os.path.join(tmp, "rtcoef.txt"), 15, 13)
        table = np.loadtxt(os.path.join(tmp, "rtcoef_test_table.txt"))
        np.savez_compressed(os.path.join(HERE, "golden_free.npz"), transform_in=tx, transform_out=ty,
                            rtcoef_in=rx, rtcoef_out=ry, rtcoef_test_table=table)
        print("golden_free.npz", tx.shape, rx.shape, table.shape)

    for cfg, (deg, n) in PLAN.items():
        with tempfile.TemporaryDirectory() as tmp:
            args = ref_configs.cmdline(cfg, n, deg, tmp)
            harness("run", args, os.path.join(tmp, "r"), tmp, {"R3D_HARNESS_TRACE": "1"})
            harness("vectors", args, tmp, tmp)
            m = FlatModel.load(os.path.join(tmp, "r.model"))
            e, c, k = load_bins(os.path.join(tmp, "r.bins"))
            fin = np.fromfile(os.path.join(tmp, "r.trace"), dtype=abi.PHONON_FINAL_DTYPE)
            assert fin.size == n
            nz = np.flatnonzero(c.sum(axis=2))            # bins that caught anything
            out = model_fields(m)
            out.update(
                run_seed=np.uint64(SEED), run_n=np.int64(n), run_counters=k, run_finals=fin,
                run_bin_index=nz.astype(np.int64), run_bin_energy=e.reshape(-1, 5)[nz], run_bin_count=c.reshape(-1, 2)[nz])
            px, py = pairs(os.path.join(tmp, "path_to_boundary.txt"), 7, 9)
            ax, ay = pairs(os.path.join(tmp, "advance.txt"), 8, 9)
            cdf = np.loadtxt(os.path.join(tmp, "cdf_search.txt"))
            out.update(path_in=px, path_out=py, advance_in=ax, advance_out=ay, cdf_cases=cdf)
            if os.path.exists(os.path.join(tmp, "catch.txt")):
                cx, cy = pairs(os.path.join(tmp, "catch.txt"), 28, 6)
                out.update(catch_in=cx, catch_out=cy)
            path = os.path.join(HERE, f"golden_{cfg}.npz")
            np.savez_compressed(path, **out)
            print(f"golden_{cfg}.npz: {os.path.getsize(path) / 1e6:.2f} MB, toa {m.n_toa}, cells {m.n_cells}, "
                  f"scat {m.n_scat}, seis {m.n_seis}, catches {int(c.sum())}, counters {k[:3]}")


if __name__ == "__main__":
    main()
