"""The oracle (oracle/r3d_oracle.c) against golden vectors made from the reference's compiled objects.

This is what pins the oracle: every fixture under tests/golden/ was written by oracle/ref_harness.cpp
calling the reference's own methods (tests/golden/make_golden.py).  The oracle keeps the reference's
operation order and is compiled without FMA contraction, so agreement is expected to the last bits;
the tolerances below (1e-12) leave room only for libm differences between machines.
"""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import CONFIGS, GOLDEN, INV_CASES, dense_bins, load_golden, rel_err
from radiative3d_b200 import abi

TOL = 1e-12   # relative; the north-star bar for deterministic sub-kernels is 1e-10


@pytest.fixture(scope="module")
def free():
    return np.load(f"{GOLDEN}/golden_free.npz")


def test_transform(free):
    out = ob.transform(free["transform_in"])
    assert rel_err(out, free["transform_out"]).max() <= TOL


def test_rtcoef_random_interfaces(free):
    out = ob.rtcoef(free["rtcoef_in"])
    ref = free["rtcoef_out"]
    assert np.array_equal(out[:, 6], ref[:, 6])                  # chosen outcome: exact
    # probabilities relative to the largest one of the row (tiny ones carry cancellation noise)
    scale = np.abs(ref[:, :6]).max(axis=1, keepdims=True)
    assert (np.abs(out[:, :6] - ref[:, :6]) / scale).max() <= TOL
    assert np.abs(out[:, 7:] - ref[:, 7:]).max() <= TOL          # unit vectors: absolute


def test_rtcoef_builtin_table(free):
    """The reference's own --rtcoef-test (rtcoef.cpp:687-742): 3 incident types x 100 angles."""
    t = free["rtcoef_test_table"]
    x = np.zeros((t.shape[0], 15))
    x[:, 2] = 1.0                                                # normal = ThetaPhi(0,0)
    x[:, 3], x[:, 5] = np.sin(t[:, 1]), np.cos(t[:, 1])          # incidence = ThetaPhi(theta,0)
    x[:, 6:12] = [10, 8, 4, 8, 4, 2]                             # rho1 a1 b1 rho2 a2 b2
    x[:, 12] = t[:, 0]
    x[:, 14] = 12345
    out = ob.rtcoef(x)
    ref = t[:, [3, 5, 7, 4, 6, 8]]                               # table order R_P T_P R_SV T_SV R_SH T_SH -> enum order
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-300)
    assert (np.abs(out[:, :6] - ref) / scale).max() <= 1e-10


@pytest.mark.parametrize("cfg", CONFIGS)
def test_path_and_advance(cfg):
    m, z = load_golden(cfg)
    out = ob.path_to_boundary(m, z["path_in"])
    ref = z["path_out"]
    assert np.array_equal(out[:, 8], ref[:, 8])                  # exit face: bit-exact
    assert rel_err(out[:, :8], ref[:, :8]).max() <= TOL
    out = ob.advance(m, z["advance_in"])
    assert rel_err(out[:, :8], z["advance_out"][:, :8]).max() <= TOL


@pytest.mark.parametrize("cfg", CONFIGS)
def test_cdf_search(cfg):
    m, z = load_golden(cfg)
    nt = m.n_toa
    for which, k, idx in z["cdf_cases"]:
        which = int(which)
        if which < 4:
            cdf = m.scat_cdf[which * nt:(which + 1) * nt]
        elif which < 6:
            cdf = m.scat_whole_cdf[(which - 4) * 4:(which - 3) * 4]
        else:
            cdf = m.src_cdf[(which - 6) * nt:(which - 5) * nt]
        assert ob.cdf_search(cdf, [int(k)])[0] == int(idx)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_catch(cfg):
    m, z = load_golden(cfg)
    out = ob.catch(m.bin_dt, m.n_bins, z["catch_in"])
    ref = z["catch_out"]
    assert np.array_equal(out[:, :2], ref[:, :2])                # caught flag and bin index: exact
    # golden energies are differences of accumulators, so compare to the bin's scale
    assert np.abs(out[:, 2:] - ref[:, 2:]).max() <= 1e-9 * max(1.0, np.abs(ref[:, 2:]).max())
    assert ref[:, 0].sum() > 0


@pytest.mark.parametrize("cfg", CONFIGS)
def test_whole_run(cfg):
    """Same Philox draws as the reference loop: end states, bins and counters must agree."""
    m, z = load_golden(cfg)
    n, seed = int(z["run_n"]), int(z["run_seed"])
    e, c, k, fin = ob.run(m, 0, n, seed, finals=True)
    ref = z["run_finals"]
    for f in ("moves", "cell", "type", "fate", "draws"):
        assert np.array_equal(fin[f], ref[f]), f
    for f in ("time", "pathlen", "amp", "theta", "phi", "pol", "loc"):
        assert rel_err(fin[f], ref[f]).max() <= TOL, f
    e_ref, c_ref = dense_bins(m, z)
    assert np.array_equal(c, c_ref)
    assert rel_err(e, e_ref).max() <= 1e-9
    assert np.array_equal(k[:3], z["run_counters"][:3])
    assert int(k[abi.R3D_CNT_PHONONS]) == n


@pytest.mark.parametrize("case", INV_CASES)
def test_invalid_phonons(case):
    """Row a22: the seven validity checks (phonons.cpp:554-584) on models the reference itself ran after the harness made
    them degenerate (tests/golden/make_inv_golden.py): mNumInvalid, mDiagInvalid and every end state, NaNs included."""
    m, z = load_golden(case, prefix="inv")
    n, seed = int(z["run_n"]), int(z["run_seed"])
    e, c, k, fin = ob.run(m, 0, n, seed, finals=True)
    ref, kr = z["run_finals"], z["run_counters"]
    assert int(kr[abi.R3D_CNT_INVALID]) > 0 and int(kr[abi.R3D_CNT_DIAG]) != 0
    assert np.array_equal(k[:3], kr[:3]) and int(k[abi.R3D_CNT_DIAG]) == int(kr[abi.R3D_CNT_DIAG])
    for f in ("moves", "cell", "type", "fate", "draws"):
        assert np.array_equal(fin[f], ref[f]), f
    for f in ("time", "pathlen", "amp", "theta", "phi", "pol", "loc"):
        assert rel_err(fin[f], ref[f]).max() <= TOL, f
    e_ref, c_ref = dense_bins(m, z)
    assert np.array_equal(c, c_ref)
    assert e.size == 0 or rel_err(e, e_ref).max() <= 1e-9          # (time_nan has no seismometers)


def test_threads_match_serial():
    m, z = load_golden("lopnor")
    e1, c1, k1, f1 = ob.run(m, 0, 600, 7, finals=True)
    e4, c4, k4, f4 = ob.run(m, 0, 600, 7, finals=True, nthreads=4)
    assert np.array_equal(c1, c4) and np.array_equal(k1, k4) and np.array_equal(f1, f4)
    assert rel_err(e1, e4).max() <= 1e-12
