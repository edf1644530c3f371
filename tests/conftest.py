import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from radiative3d_b200.model import FlatModel, _ARRAYS  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CONFIGS = ["halfspace", "halfspace_nearsrc50", "crustpinch", "lopnor", "spherical"]
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_harness")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


INV_CASES = ["path_nan", "time_nan", "path_negative", "stuck", "slow", "loop_exceed"]     # tests/golden/make_inv_golden.py


def load_golden(cfg, prefix="golden"):
    """(FlatModel, npz dict) of tests/golden/<prefix>_<cfg>.npz (made from the reference by make_golden.py / make_inv_golden.py)."""
    z = np.load(os.path.join(GOLDEN, f"{prefix}_{cfg}.npz"))
    f, i = z["scalars_f"], z["scalars_i"]
    m = FlatModel(freq_hz=float(f[0]), ttl=float(f[1]), bin_dt=float(f[2]), earth_center=tuple(map(float, f[3:6])),
                  min_theta=float(f[6]), max_theta=float(f[7]), slow_concern=float(f[8]), src_loc=tuple(map(float, f[9:12])),
                  cyl_radius2=float(f[12]), loop_concern=int(i[0]), n_bins=int(i[1]), ecs_radial=int(i[2]),
                  no_deflect=int(i[3]), src_cell=int(i[4]), cell_kind=int(i[5]))
    for name, dt in _ARRAYS:
        setattr(m, name, np.ascontiguousarray(z[name], dtype=dt))
    return m.validate(), z


def dense_bins(model, z):
    e = np.zeros((model.n_seis * model.n_bins, 5))
    c = np.zeros((model.n_seis * model.n_bins, 2), dtype=np.uint64)
    e[z["run_bin_index"]] = z["run_bin_energy"]
    c[z["run_bin_index"]] = z["run_bin_count"]
    return e.reshape(model.n_seis, model.n_bins, 5), c.reshape(model.n_seis, model.n_bins, 2)


def rel_err(a, b):
    """Element-wise relative error with matching infinities / NaNs counted as equal."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return np.where(same, 0.0, np.where(np.isfinite(err), err, np.inf))


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
