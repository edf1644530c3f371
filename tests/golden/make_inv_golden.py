#!/usr/bin/env python
"""Generate tests/golden/inv_<case>.npz: invalid-phonon fixtures (reference phonons.cpp:554-584,
dataout.cpp:611-617), decided by the UNMODIFIED reference.

No stock model produces a single INV phonon, so each case takes a BASELINE model the reference built, overwrites members of
the reference's own objects (oracle/ref_harness.cpp, R3D_HARNESS_MUTATE) so that one of the seven validity checks must
fire, flattens THAT model and runs the reference's GenerateEventPhonon()+Propagate() loop on the Philox draw stream.  The
reference therefore decides mNumInvalid, mDiagInvalid and every phonon's end state; oracle and GPU are held to them.

    python tests/golden/make_inv_golden.py

case            model       mutation                          reason(s) the reference reports
path_nan        spherical   SphereShell::mZeroRadius2[P]=NaN  INV_PATH_NAN: every P arc is NaN, the ray "reflects" at the first
                                                              discontinuity below it for ever (R/T default choice, rtcoef.cpp:459)
time_nan        halfspace   mVelTop[S]=NaN, S MFP 3 km        INV_TIME_NAN (no seismometers: the reference converts a NaN arrival
                                                              time to an unsigned bin index, dataout.cpp:163 - undefined behaviour)
path_negative   halfspace   S MFP -3 km                       INV_PATH_NEGATIVE, and INV_TIME_NEGATIVE for phonons whose long P legs
                                                              keep the path length positive while the recent travel time is negative
stuck           halfspace   S MFP 0, free surface absorbing   INV_STUCK (with the surface reflecting, P->SV conversions put S phonons ON the
                                                              surface plane, where "0 < distance to the face" is decided by the last bit
                                                              of the position: the outcome would hinge on FMA contraction, not on the check)
slow            halfspace   S MFP 3 km, cm_slow_concern 1e4 s INV_SLOW
loop_exceed     halfspace   S MFP 3 km, cm_loop_concern 200,  INV_LOOP_EXCEED
                            TTL 1e9 s
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from radiative3d_b200 import workloads, abi  # noqa: E402
from radiative3d_b200.model import FlatModel  # noqa: E402
from oracle_binding import load_bins  # noqa: E402
from make_golden import harness, model_fields, SEED, HARNESS  # noqa: E402
from make_event_golden import parse_reports  # noqa: E402

# case -> (model, toa degree, phonons, mutation, keep seismometers, expected diag mask)
CASES = {
    "path_nan": ("spherical", 2, 400, "shell_zr2_p=nan", True, 0x01),
    "time_nan": ("halfspace", 2, 1500, "cyl_vel_s=nan,mfp_s=3", False, 0x02),
    "path_negative": ("halfspace", 2, 1500, "mfp_s=-3", True, 0x0c),
    "stuck": ("halfspace", 2, 1500, "mfp_s=0,no_reflect=1", True, 0x10),
    "slow": ("halfspace", 2, 1500, "mfp_s=3,slow=1e4", True, 0x20),
    "loop_exceed": ("halfspace", 2, 1500, "mfp_s=3,loop=200,ttl=1e9", True, 0x40),
}


def main():
    if not os.path.exists(HARNESS):
        sys.exit("oracle/_ref/r3d_ref_harness is missing: run `make -C oracle ref` where /root/reference exists")
    for case, (cfg, deg, n, mut, seis, diag) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            args = [a for a in workloads.cmdline(cfg, n, deg, tmp) if seis or not a.startswith("--seis")]
            harness("run", args, os.path.join(tmp, "r"), tmp, {"R3D_HARNESS_TRACE": "1", "R3D_HARNESS_MUTATE": mut})
            m = FlatModel.load(os.path.join(tmp, "r.model"))
            e, c, k = load_bins(os.path.join(tmp, "r.bins"))
            fin = np.fromfile(os.path.join(tmp, "r.trace"), dtype=abi.PHONON_FINAL_DTYPE)
            # the reference's own INV report lines (--reports=INV is part of every BASELINE command line)
            ev = parse_reports(os.path.join(tmp, "reports.dat"))
        assert fin.size == n and int(k[7]) == diag, (case, k)
        assert int((ev["kind"] == 7).sum()) == int(k[2])
        nz = np.flatnonzero(c.sum(axis=2))
        out = model_fields(m)
        out.update(run_seed=np.uint64(SEED), run_n=np.int64(n), run_counters=k, run_finals=fin, run_bin_index=nz.astype(np.int64),
                   run_bin_energy=e.reshape(-1, 5)[nz], run_bin_count=c.reshape(-1, 2)[nz], mutation=np.array(mut),
                   inv_lines=ev[ev["kind"] == 7])
        path = os.path.join(HERE, f"inv_{case}.npz")
        np.savez_compressed(path, **out)
        reasons = sorted({int(f) >> 8 for f in fin["fate"] if (int(f) & 0xFF) == abi.R3D_FATE_INVALID})
        print(f"inv_{case}.npz: {os.path.getsize(path) / 1e3:.0f} kB, {cfg} [{mut}], lost/timeout/invalid {k[:3]}, diag 0x{int(k[7]):02x}, "
              f"per-phonon reason masks {[hex(r) for r in reasons]}")


if __name__ == "__main__":
    main()
