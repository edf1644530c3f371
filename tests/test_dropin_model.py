"""The model the drop-in program builds (integration/_build/r3d_gpu_main: the reference's own main(), command line, model
plugins and Scatterer bookkeeping, with Scatterer::PopulateProbDists and the phonon loop replaced) against the model the
UNMODIFIED reference builds (oracle/_ref/r3d_ref_harness, mode dump).

  * not gpu: with the reference's own GSATO loop (R3D_GPU_SCATTERERS=0) the two flattened models are identical bit for bit,
    which pins the symbol replacement (weakened Scatterer::PopulateProbDists, wrapped main) and the flattener; the 64-bit
    --num-phonons reaches out_mparams.octv.
  * gpu: with the G values from the device, for all five BASELINE models: CDF values within 1e-10 of the
    reference's (relative to each table's total), every other array identical, and 1e6 scripted draws pick the same table
    index in both sets of tables (ProbDist::GetRandomIndex, probability.cpp:104-129).
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_binding as ob
from conftest import CONFIGS, REF_HARNESS
from radiative3d_b200 import reference_host, workloads
from radiative3d_b200.model import FlatModel, _ARRAYS

TABLES = ("scat_cdf", "scat_spol", "scat_whole_cdf", "scat_mfp")


def need_binaries():
    if not os.path.exists(reference_host.GPU_MAIN) or not os.path.exists(REF_HARNESS):
        pytest.skip("integration/_build/r3d_gpu_main or oracle/_ref/r3d_ref_harness was not built (needs the reference checkout)")


def reference_model(cfg, deg):
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, R3D_HARNESS="dump", R3D_HARNESS_OUT=os.path.join(tmp, "m"))
        p = subprocess.run([REF_HARNESS] + workloads.cmdline(cfg, 10, deg, tmp), cwd=tmp, env=env, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-1500:]
        return FlatModel.load(os.path.join(tmp, "m"))


@pytest.mark.parametrize("cfg,deg", [("halfspace", 4), ("crustpinch", 3), ("lopnor", 4), ("spherical", 4)])
def test_cpu_tables_reproduce_the_reference_model(cfg, deg):
    need_binaries()
    m, r = reference_host.build_model(cfg, deg, gpu_tables=False), reference_model(cfg, deg)
    for name, _ in _ARRAYS:
        assert np.array_equal(getattr(m, name), getattr(r, name)), name
    assert (m.n_toa, m.n_cells, m.n_scat, m.n_seis, m.n_bins, m.ttl, m.src_cell) == (r.n_toa, r.n_cells, r.n_scat, r.n_seis, r.n_bins, r.ttl, r.src_cell)


def test_64_bit_phonon_count_reaches_the_parameter_file():
    need_binaries()
    with tempfile.TemporaryDirectory() as tmp:
        args = [a for a in workloads.cmdline("halfspace", 10, 2, tmp) if not a.startswith("--num-phonons")]
        args += ["--num-phonons=10B", "--mparams-outfile=out_mparams.octv", "--seed=5", "--gpu-devices=0,1"]
        env = dict(os.environ, R3D_GPU_SCATTERERS="0", R3D_GPU_DUMP_MODEL=os.path.join(tmp, "m"), R3D_GPU_DUMP_ONLY="1")
        p = subprocess.run([reference_host.GPU_MAIN] + args, cwd=tmp, env=env, capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
        txt = open(os.path.join(tmp, "out_mparams.octv")).read()
    assert "# name: NumPhonons \n# type: scalar \n10000000000 " in txt
    # an option the reference does not know is still the reference's error
    with tempfile.TemporaryDirectory() as tmp:
        p = subprocess.run([reference_host.GPU_MAIN] + workloads.cmdline("halfspace", 10, 2, tmp) + ["--num-phonons=12Q"], cwd=tmp,
                           env=dict(os.environ, R3D_GPU_SCATTERERS="0"), capture_output=True, text=True)
        assert p.returncode == 1 and "Error processing command-line option" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS)
def test_gpu_tables_match_the_reference_build(cfg):
    need_binaries()
    deg = 6                                            # 81 920 take-off angles
    g, r = reference_host.build_model(cfg, deg, gpu_tables=True), reference_model(cfg, deg)
    nt, ns = r.n_toa, r.n_scat
    for name, _ in _ARRAYS:
        if name not in TABLES:
            assert np.array_equal(getattr(g, name), getattr(r, name)), name
    cg, cr = g.scat_cdf.reshape(ns * 4, nt), r.scat_cdf.reshape(ns * 4, nt)
    scale = cr[:, -1:].copy()
    scale[scale == 0] = 1.0
    assert (np.abs(cg - cr) / scale).max() <= 1e-10
    assert np.abs(g.scat_mfp / r.scat_mfp - 1).max() <= 1e-10
    assert np.abs(g.scat_whole_cdf - r.scat_whole_cdf).max() <= 1e-10 * np.abs(r.scat_whole_cdf).max()
    # 1e6 scripted draws, spread over every table of the model: same index from both builds
    rng = np.random.default_rng(2026)
    per = max(1, 1_000_000 // (ns * 4))
    differ = total = 0
    for t in range(ns * 4):
        if cr[t, -1] == 0:
            continue
        k = rng.integers(0, 2**31, per, dtype=np.uint32)
        k[:3] = (0, 1, 2**31 - 1)
        differ += int((ob.cdf_search(cg[t], k) != ob.cdf_search(cr[t], k)).sum())
        total += per
    assert differ == 0, f"{differ} of {total} draws pick another index"
