"""Statistical parity of whole runs with the reference program's OWN output (BASELINE.json north_star).

tests/golden/stat_<cfg>.npz hold runs of the unmodified reference binary with its own rand() stream (make_stat_golden.py):
P independent processes of n phonons each; per process the window-summed seismometer counts and energies, and the loss counters.
The GPU traces P batches of n phonons of the same model with the Philox stream, so for every (seismometer, time window) cell
there are two independent samples of P batch values of the same physics.  Catches are not Poisson (a reverberating phonon is
caught many times in one cell) and energies are heavy-tailed, so the Monte-Carlo error is taken from the batch-to-batch scatter
of BOTH samples (Welch):

    t = (mean_g - mean_c) / sqrt(var_g / P + var_c / P)        per cell, for counts and for energies

over the cells that hold at least 100 catches in the two samples together.  Criteria: rms(t) <= 1.5 (1.0-1.25 seen reference vs reference),
max |t| <= 7, the same t for the batch TOTALS (all cells summed: a common bias; cells are strongly correlated through phonons
caught many times, so the bias is tested on the sum, not on the cell statistics) within 5, and the lost / time-out fractions
within 5 sqrt(2 p (1 - p) / N).

The same criteria applied to the reference against itself (first half of the processes vs the second half) run on the CPU in
the not-gpu suite, which calibrates them; a GPU run with the mean free paths scaled by 1.3 must FAIL them (the test has power).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

CONFIGS = ["halfspace", "halfspace_nearsrc50", "crustpinch", "lopnor", "spherical"]


def windows(a, w):
    n = a.shape[1]
    edges = [n * i // w for i in range(w + 1)]
    return np.stack([a[:, edges[i]:edges[i + 1]].sum(axis=1) for i in range(w)], axis=1)


def load_stat(cfg):
    path = os.path.join(GOLDEN, f"stat_{cfg}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


def welch(xa, xb, sel):
    """t statistic per selected cell for two samples of batches [Pa, ...] and [Pb, ...]."""
    pa, pb = xa.shape[0], xb.shape[0]
    den = np.sqrt(xa.var(axis=0, ddof=1) / pa + xb.var(axis=0, ddof=1) / pb)
    ok = sel & (den > 0)
    return ((xa.mean(axis=0) - xb.mean(axis=0))[ok] / den[ok])


def compare(counts_a, counts_b, energy_a, energy_b, counters_a, counters_b, n_a, n_b):
    """counts_* [P, n_seis, W, 2] and energy_* [P, n_seis, W] per batch; counters_* summed [lost, timeout, ...]; n_* phonons
    in each sample.  Returns (statistics, failures)."""
    fails, stats = [], {}
    ca, cb = counts_a.sum(-1).astype(np.float64), counts_b.sum(-1).astype(np.float64)
    sel = (ca.sum(0) + cb.sum(0)) >= 100
    for name, xa, xb in (("counts", ca, cb), ("energy", energy_a, energy_b)):
        t = welch(xa, xb, sel)
        stats[name] = {"cells": int(t.size), "rms_t": float(np.sqrt((t ** 2).mean())) if t.size else 0.0,
                       "max_t": float(np.abs(t).max()) if t.size else 0.0,
                       "t_total": float(welch(xa.reshape(xa.shape[0], -1).sum(1)[:, None], xb.reshape(xb.shape[0], -1).sum(1)[:, None],
                                              np.array([True]))[0])}
        if t.size < 20:
            fails.append(f"{name}: too few populated cells ({t.size})")
            continue
        if stats[name]["rms_t"] > 1.5:
            fails.append(f"{name}: rms t = {stats[name]['rms_t']:.2f} over {t.size} cells")
        if stats[name]["max_t"] > 7.0:
            fails.append(f"{name}: max |t| = {stats[name]['max_t']:.2f}")
        if abs(stats[name]["t_total"]) > 5.0:
            fails.append(f"{name}: batch totals differ, t = {stats[name]['t_total']:.2f}")
    for i, name in enumerate(("lost", "timeout")):
        pa, pb = counters_a[i] / n_a, counters_b[i] / n_b
        p = (counters_a[i] + counters_b[i]) / (n_a + n_b)
        if abs(pa - pb) > 5.0 * np.sqrt(p * (1.0 - p) * (1.0 / n_a + 1.0 / n_b)) + 1e-12:
            fails.append(f"{name} fraction {pa:.5f} vs {pb:.5f}")
    return stats, fails


@pytest.mark.parametrize("cfg", CONFIGS)
def test_reference_halves_agree(cfg):
    """Calibration on the CPU: the reference's first P/2 processes against its last P/2 must pass the criteria."""
    z = load_stat(cfg)
    c, e, k = z["counts"], z["energy"].astype(np.float64).sum(-1), z["counters"]
    h = c.shape[0] // 2
    n = int(z["n_each"]) * h
    stats, fails = compare(c[:h], c[h:], e[:h], e[h:], k[:h].sum(0), k[h:].sum(0), n, n)
    assert not fails, (stats, fails)


def gpu_batches(m, z, seed=777):
    """The GPU's P batches of n phonons each (consecutive index ranges of one seed)."""
    from radiative3d_b200 import engine
    p, n_each, w = z["counts"].shape[0], int(z["n_each"]), int(z["windows"])
    cs, es, ks = [], [], np.zeros(3, dtype=np.int64)
    with engine.Engine(m) as eng:
        for b in range(p):
            eng.reset()
            eng.run_simulation(n_each, seed=seed, first_phonon=b * n_each)
            e, c, k = eng.fetch()
            cs.append(windows(c.astype(np.int64), w))
            es.append(windows(e[..., 3:5], w).sum(-1))
            ks += k[:3].astype(np.int64)
    return np.stack(cs), np.stack(es), ks, p * n_each


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS)
def test_gpu_matches_reference_statistics(cfg):
    z = load_stat(cfg)
    m, _ = load_golden(cfg)
    cg, eg, kg, n = gpu_batches(m, z)
    c, e, k = z["counts"], z["energy"].astype(np.float64).sum(-1), z["counters"]
    stats, fails = compare(cg, c, eg, e, kg, k.sum(0), n, n)
    print(cfg, stats)
    assert not fails, (stats, fails)


@pytest.mark.gpu
def test_statistics_detect_wrong_physics():
    """Power of the test: 30 % longer mean free paths must be rejected."""
    z = load_stat("halfspace")
    m, _ = load_golden("halfspace")
    m.scat_mfp = m.scat_mfp * 1.3
    cg, eg, kg, n = gpu_batches(m, z)
    c, e, k = z["counts"], z["energy"].astype(np.float64).sum(-1), z["counters"]
    stats, fails = compare(cg, c, eg, e, kg, k.sum(0), n, n)
    assert fails, stats
