"""Event reports (DataReporter::Report*, dataout.cpp:484-617) against the reference's own report lines.

tests/golden/events_<cfg>.npz hold the reference's reports.dat (--reports=ALL_ON) for its loop run on the Philox draw stream
(make_event_golden.py).  The GPU traces the same phonons, so the two event streams must agree line by line: kind, ray type and
move count exactly; time, path length, position, direction and amplitude to the 6 digits the reference prints.
  * through the drop-in program (reference main + its own Report* printers fed from r3d_trace_events), all five models;
  * through the C ABI directly (r3d_trace_events), where the output coordinates equal the model coordinates.
This is also the finest-grained parity check of the propagate path as a whole: every event of every phonon.
"""
import os
import sys

import numpy as np
import pytest

from conftest import CONFIGS, GOLDEN, load_golden
from radiative3d_b200 import abi, engine, reference_host

sys.path.insert(0, GOLDEN)
from make_event_golden import parse_reports, report_args  # noqa: E402

pytestmark = pytest.mark.gpu


def unit(theta, phi):
    return np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], -1)


def compare_streams(kind, typ, it, phonon, time, pathlen, xyz, theta, phi, amp, ref, min_same=0.95):
    """Phonon by phonon: discrete columns exact and numeric columns to print precision for phonons whose event sequences
    match; at least `min_same` of the phonons must match (a trajectory that diverges by rounding shows up as a different
    sequence, exactly as in test_gpu_propagate)."""
    n = int(ref["phonon"].max()) + 1
    same = 0
    scale = max(1.0, np.abs(ref["xyz"]).max())
    for p in range(n):
        a, b = np.flatnonzero(phonon == p), np.flatnonzero(ref["phonon"] == p)
        if a.size != b.size or not (np.array_equal(kind[a], ref["kind"][b]) and np.array_equal(typ[a], ref["type"][b]) and np.array_equal(it[a], ref["it"][b])):
            continue
        same += 1
        assert np.allclose(time[a], ref["time"][b], rtol=3e-5, atol=1e-9), p
        assert np.allclose(pathlen[a], ref["pathlen"][b], rtol=3e-5, atol=1e-9), p
        assert np.abs(xyz[a] - ref["xyz"][b]).max() <= 3e-5 * scale, p
        assert np.abs(unit(theta[a], phi[a]) - unit(ref["theta"][b], ref["phi"][b])).max() <= 3e-5, p
        assert np.allclose(amp[a], ref["amp"][b], rtol=3e-5, atol=1e-12), p
    assert same >= min_same * n, f"only {same} of {n} phonons have the reference's event sequence"
    return same, n


@pytest.mark.parametrize("cfg", CONFIGS)
def test_report_file_of_the_drop_in_program(cfg, tmp_path):
    if not os.path.exists(reference_host.GPU_MAIN):
        pytest.skip("integration/_build/r3d_gpu_main was not built (needs the reference checkout at build time)")
    z = np.load(os.path.join(GOLDEN, f"events_{cfg}.npz"))
    ref, n, seed, deg = z["events"], int(z["n"]), int(z["seed"]), int(z["toa_degree"])
    os.makedirs(tmp_path, exist_ok=True)
    import subprocess
    env = dict(os.environ, R3D_GPU_SEED=str(seed))
    p = subprocess.run([reference_host.GPU_MAIN] + report_args(cfg, n, deg, str(tmp_path)), cwd=str(tmp_path), env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    ev = parse_reports(os.path.join(tmp_path, "reports.dat"))
    same, tot = compare_streams(ev["kind"], ev["type"], ev["it"], ev["phonon"], ev["time"], ev["pathlen"], ev["xyz"], ev["theta"], ev["phi"], ev["amp"], ref)
    print(cfg, f"{same}/{tot} phonons identical event sequences, {ev.size} lines")


@pytest.mark.parametrize("cfg", ["halfspace", "halfspace_nearsrc50"])
def test_events_through_the_abi(cfg):
    z = np.load(os.path.join(GOLDEN, f"events_{cfg}.npz"))
    ref, n, seed = z["events"], int(z["n"]), int(z["seed"])
    m, _ = load_golden(cfg)
    with engine.Engine(m) as eng:
        ev = eng.trace_events(n, seed=seed)
        e, c, k = eng.fetch()
        only = eng.trace_events(n, seed=seed, kinds=(1 << abi.R3D_EV_SCT) | (1 << abi.R3D_EV_LST))
    assert int(k[abi.R3D_CNT_PHONONS]) == n
    # sorted by (phonon, seq), seq counting from 0 within each phonon
    assert np.all(np.diff(ev["phonon"].astype(np.int64)) >= 0)
    first = np.flatnonzero(np.r_[True, np.diff(ev["phonon"].astype(np.int64)) > 0])
    assert np.all(ev["seq"][first] == 0) and np.all(ev["kind"][first] == abi.R3D_EV_GEN)
    compare_streams(ev["kind"], ev["type"], ev["moves"], ev["phonon"].astype(np.int64), ev["time"], ev["pathlen"], ev["loc"], ev["theta"], ev["phi"], ev["amp"], ref)
    # the mask selects kinds; the tallies say the same as the stream
    assert set(np.unique(only["kind"])) <= {abi.R3D_EV_SCT, abi.R3D_EV_LST}
    assert int((ev["kind"] == abi.R3D_EV_SCT).sum()) == int(k[abi.R3D_CNT_SCATTERS]) == int((only["kind"] == abi.R3D_EV_SCT).sum())
    assert int((ev["kind"] == abi.R3D_EV_LST).sum()) == int(k[abi.R3D_CNT_LOST])
    assert int((ev["kind"] == abi.R3D_EV_TMO).sum()) == int(k[abi.R3D_CNT_TIMEOUT])


def test_event_buffer_too_small():
    m, _ = load_golden("halfspace")
    with engine.Engine(m) as eng:
        with pytest.raises(engine.R3DError):
            eng.trace_events(200, seed=1, capacity=50)
