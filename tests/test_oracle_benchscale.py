"""The oracle against the reference AT the benchmarked configuration (take-off-angle degree 9: 5 242 880 angles).

The committed fixtures pin the oracle at degrees 2-3 (small files); tests/test_gpu_benchscale.py holds the GPU to the
oracle at degree 9.  This test closes the chain where the reference checkout is built (the build container): the
reference's own GenerateEventPhonon() + Propagate() loop, driven by the Philox stream through oracle/_ref/r3d_ref_harness at
degree 9, against the oracle on the model that run flattened - end states at 1e-12, discrete outcomes, bins and counters
exactly, like tests/test_oracle_golden.py::test_whole_run."""
import os
import tempfile

import numpy as np
import pytest

import oracle_binding as ob
from conftest import REF_HARNESS, rel_err
from radiative3d_b200 import abi, workloads
from radiative3d_b200.model import FlatModel
from golden.make_golden import SEED, harness

TOL = 1e-12


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref/r3d_ref_harness not built (needs the reference checkout)")
@pytest.mark.parametrize("cfg,n", [("halfspace_nearsrc50", 30000), ("lopnor", 1500)])
def test_oracle_matches_reference_at_degree_9(cfg, n):
    if cfg == "lopnor" and not os.environ.get("R3D_SLOW_TESTS"):
        pytest.skip("3.5 GB of tables and ~1 min of reference model build: set R3D_SLOW_TESTS=1")
    with tempfile.TemporaryDirectory() as tmp:
        harness("run", workloads.cmdline(cfg, n, 9, tmp), os.path.join(tmp, "r"), tmp, {"R3D_HARNESS_TRACE": "1"})
        m = FlatModel.load(os.path.join(tmp, "r.model"))
        e_ref, c_ref, k_ref = ob.load_bins(os.path.join(tmp, "r.bins"))
        ref = np.fromfile(os.path.join(tmp, "r.trace"), dtype=abi.PHONON_FINAL_DTYPE)
    assert m.n_toa == 5242880 and ref.size == n
    e, c, k, fin = ob.run(m, 0, n, SEED, finals=True, nthreads=min(8, os.cpu_count() or 1))
    for f in ("moves", "cell", "type", "fate", "draws"):
        assert np.array_equal(fin[f], ref[f]), f
    for f in ("time", "pathlen", "amp", "theta", "phi", "pol", "loc"):
        assert rel_err(fin[f], ref[f]).max() <= TOL, f
    assert np.array_equal(c, c_ref) and np.array_equal(k[:3], k_ref[:3])
    assert rel_err(e, e_ref).max() <= 1e-9
