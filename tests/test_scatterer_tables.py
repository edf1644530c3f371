"""Scatterer-table construction (Scatterer ctor, scatterers.cpp:134-220 + scatparams.cpp:75-194).

tests/golden/scat_params.npz holds the ScatterParams of every scatterer of the golden models, golden_<cfg>.npz the tables the
reference built from them.  The oracle's restatement is held to those tables on the CPU; the device version (G values on the
GPU, cumulative sums on the host in index order) to the oracle and the golden tables on the GPU: 1e-10 relative (north_star).
"""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import CONFIGS, GOLDEN, load_golden


def golden_tables(cfg):
    m, _ = load_golden(cfg)
    pars = np.load(f"{GOLDEN}/scat_params.npz")[cfg]
    nt = m.n_toa
    assert pars.shape == (m.n_scat, 6)
    for i in range(m.n_scat):
        yield (pars[i], m.toa_theta, m.toa_phi, m.scat_cdf[i * 4 * nt:(i + 1) * 4 * nt].reshape(4, nt), m.scat_spol[i * nt:(i + 1) * nt],
               m.scat_whole_cdf[i * 8:(i + 1) * 8].reshape(2, 4), m.scat_mfp[i * 2:(i + 1) * 2])


def check(built, ref, tol):
    cdf, spol, whole, mfp = built
    rcdf, rspol, rwhole, rmfp = ref
    scale = rcdf[:, -1:].copy()
    scale[scale == 0] = 1.0
    assert (np.abs(cdf - rcdf) / scale).max() <= tol                       # cumulative sums relative to their totals
    assert np.abs(whole - rwhole).max() <= tol * max(1e-300, np.abs(rwhole).max())
    assert np.abs(mfp / rmfp - 1.0).max() <= tol
    # the polarisation angle is atan2 of two amplitudes: compare where it is defined (S->S weight not negligible)
    w = np.diff(np.r_[0.0, rcdf[3]])
    ok = w > 1e-12 * w.max()
    d = np.abs(np.angle(np.exp(1j * (spol - rspol))))
    assert d[ok].max() <= max(tol, 1e-9)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_oracle_tables_match_reference(cfg):
    for par, th, ph, *ref in golden_tables(cfg):
        check(ob.build_scatterer_tables(par, th, ph), ref, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS)
def test_device_tables_match_reference(cfg):
    from radiative3d_b200 import engine
    for par, th, ph, *ref in golden_tables(cfg):
        built = engine.build_scatterer_tables(par, th, ph)
        check(built, ref, 1e-10)
        check(built, ob.build_scatterer_tables(par, th, ph), 1e-10)


@pytest.mark.gpu
def test_device_tables_full_size_and_draws():
    """TOA-degree-9 size (5 242 880 angles on a Fibonacci-like set): against the oracle, and table look-ups on the two
    sets of CDFs pick the same index for all but a vanishing share of draws (an entry boundary within rounding of r)."""
    from radiative3d_b200 import engine
    n = 5242880
    k = np.arange(n) + 0.5
    th, ph = np.arccos(1 - 2 * k / n), (np.pi * (1 + 5 ** 0.5) * k) % (2 * np.pi) - np.pi
    par = np.load(f"{GOLDEN}/scat_params.npz")["halfspace"][0]
    dev, ref = engine.build_scatterer_tables(par, th, ph), ob.build_scatterer_tables(par, th, ph)
    check(dev, ref, 1e-10)
    draws = np.random.default_rng(5).integers(0, 2**31, 200000, dtype=np.uint32)
    for t in range(4):
        a, b = engine.cdf_search(dev[0][t], draws), ob.cdf_search(ref[0][t], draws)
        assert (a != b).mean() <= 1e-4


@pytest.mark.gpu
def test_batch_builds_a_models_tables():
    """All 21 scatterers of the Lop Nor model in one call, laid out as r3d_model_desc wants them."""
    from radiative3d_b200 import engine
    m, _ = load_golden("lopnor")
    pars = np.load(f"{GOLDEN}/scat_params.npz")["lopnor"]
    cdf, spol, whole, mfp = engine.build_scatterer_tables(pars, m.toa_theta, m.toa_phi)
    scale = np.abs(m.scat_cdf).max()
    assert np.abs(cdf.ravel() - m.scat_cdf).max() <= 1e-10 * scale
    assert np.abs(mfp.ravel() / m.scat_mfp - 1).max() <= 1e-10
    assert np.abs(whole.ravel() - m.scat_whole_cdf).max() <= 1e-10 * np.abs(m.scat_whole_cdf).max()
