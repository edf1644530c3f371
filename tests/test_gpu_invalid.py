"""Row a22 on the device: the seven validity checks of Phonon::Propagate (phonons.cpp:554-584) and the counters /
diagnostic mask of DataReporter::ReportInvalidPhonon (dataout.cpp:611-617).

No stock model produces an invalid phonon, so tests/golden/inv_<case>.npz hold models the reference itself ran after the
harness overwrote members of its objects (NaN / negative / zero mean free paths, NaN velocities, NaN arc parameters, tiny
loop and huge slow thresholds; tests/golden/make_inv_golden.py): mNumInvalid, mDiagInvalid, the INV report lines and every
phonon's end state are the reference's.  The kernel traces the same phonons on the same draw stream.
"""
import numpy as np
import pytest

from conftest import INV_CASES, dense_bins, load_golden, rel_err
from radiative3d_b200 import abi, engine

pytestmark = pytest.mark.gpu

REASON_BIT = {"path_nan": 0x01, "time_nan": 0x02, "path_negative": 0x04 | 0x08, "stuck": 0x10, "slow": 0x20, "loop_exceed": 0x40}


@pytest.mark.parametrize("case", INV_CASES)
def test_invalid_phonons_match_reference(case):
    m, z = load_golden(case, prefix="inv")
    n, seed = int(z["run_n"]), int(z["run_seed"])
    ref, kr = z["run_finals"], z["run_counters"]
    with engine.Engine(m) as eng:
        fin = eng.trace(n, seed)
        e, c, k = eng.fetch()
    diag = int(k[abi.R3D_CNT_DIAG])
    # discrete outcome, INVALID reason mask included (fate = R3D_FATE_INVALID | mask << 8): identical for >= 99.5 % ...
    same = np.ones(n, dtype=bool)
    for f in ("moves", "cell", "type", "fate", "draws"):
        same &= fin[f] == ref[f]
    assert same.mean() >= 0.995, f"only {same.mean():.4f} of phonons share the reference's outcome"
    # ... and the numbers (NaN == NaN, -inf == -inf) to 1e-8 for those
    for f in ("time", "pathlen"):
        assert rel_err(fin[f][same], ref[f][same]).max() <= 1e-8, f
    n_off = int((~same).sum())
    assert abs(int(k[abi.R3D_CNT_INVALID]) - int(kr[abi.R3D_CNT_INVALID])) <= n_off
    assert int(k[abi.R3D_CNT_INVALID]) > 0
    assert int(k[:3].sum()) == n
    # the diagnostic word is the OR of the reasons: exactly the reference's mDiagInvalid
    assert diag == int(kr[abi.R3D_CNT_DIAG]) == REASON_BIT[case]
    if n_off == 0:
        assert np.array_equal(k[:3], kr[:3])
        e_ref, c_ref = dense_bins(m, z)
        assert np.array_equal(c, c_ref)
    # per-phonon reason masks are the reference's
    inv = (ref["fate"] & 0xFF) == abi.R3D_FATE_INVALID
    assert np.array_equal(fin["fate"][same & inv] >> 8, ref["fate"][same & inv] >> 8)


@pytest.mark.parametrize("case", ["path_negative", "loop_exceed"])
def test_inv_events_carry_the_reason(case):
    """r3d_trace_events: one INV record per invalid phonon, at the reference's INV report line, with the reason mask."""
    m, z = load_golden(case, prefix="inv")
    n, seed = int(z["run_n"]), int(z["run_seed"])
    lines = z["inv_lines"]                    # the reference's own "INV:" lines of the same run (reports.dat)
    with engine.Engine(m) as eng:
        ev = eng.trace_events(n, seed=seed, kinds=1 << abi.R3D_EV_INV)
        e, c, k = eng.fetch()
    assert ev.size == int(k[abi.R3D_CNT_INVALID]) == lines.size
    assert np.all(ev["kind"] == abi.R3D_EV_INV)
    ref = z["run_finals"]
    want = ref["fate"][ev["phonon"].astype(np.int64)]
    assert np.all((want & 0xFF) == abi.R3D_FATE_INVALID)
    assert np.array_equal(ev["reason"], want >> 8)
    assert int(np.bitwise_or.reduce(ev["reason"])) == int(z["run_counters"][abi.R3D_CNT_DIAG])
    # same lines as the reference printed (6 digits): move count exactly, time and path length to print precision
    assert np.array_equal(ev["moves"], lines["it"])
    assert np.allclose(ev["time"], lines["time"], rtol=3e-5, atol=1e-9, equal_nan=True)
    assert np.allclose(ev["pathlen"], lines["pathlen"], rtol=3e-5, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("case", INV_CASES)
def test_plain_and_trace_kernels_count_the_same_invalid_phonons(case):
    """The plain kernels keep the path length and the recent travel time - the two sums that are only ever TESTED by the
    validity checks (NaN, sign, zero, below the slow-concern threshold) - as FP32; the trace kernels, which the reference
    fixtures above are compared with, as FP64.  Both must invalidate the same phonons for the same reasons."""
    m, z = load_golden(case, prefix="inv")
    n, seed = int(z["run_n"]), int(z["run_seed"])
    with engine.Engine(m) as eng:
        eng.trace(n, seed)
        _, c_t, k_t = eng.fetch()
        eng.reset()
        eng.run_simulation(n, seed=seed)
        _, c_p, k_p = eng.fetch()
    assert np.array_equal(k_t[:3], k_p[:3]) and int(k_t[abi.R3D_CNT_DIAG]) == int(k_p[abi.R3D_CNT_DIAG])
    assert np.array_equal(c_t, c_p) and int(k_t[abi.R3D_CNT_EVENTS]) == int(k_p[abi.R3D_CNT_EVENTS])
