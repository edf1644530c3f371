"""The propagate kernel (through r3d_create / r3d_run / r3d_trace / r3d_fetch) against the reference's own loop.

tests/golden holds end states, bins and counters of the reference's GenerateEventPhonon()+Propagate() run on the
Philox draw stream; the kernel consumes the same stream, so phonons can be compared one by one.  FP64 rounding
differs (FMA contraction, CUDA libm), and a trajectory is a chain of up to hundreds of dependent events, so:
  * discrete outcome (fate, cell, type, move count, draws used) must match for >= 99.5 % of phonons,
  * for those, time / path length / amplitude / location must match to 1e-8 relative,
  * bin counts must match except for the phonons above; energies to 1e-8 of the bin.
"""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import CONFIGS, REF_HARNESS, dense_bins, load_golden, rel_err
from radiative3d_b200 import abi, engine

pytestmark = pytest.mark.gpu


def bin_energy_err(e, e_ref):
    """Largest |difference| of any energy channel relative to its bin's total energy (axis components of a bin
    can be exactly 0 in the reference and 1e-35 on the device)."""
    scale = np.maximum(e_ref[..., 3:].sum(axis=-1, keepdims=True), 1e-300)
    return (np.abs(e - e_ref) / scale).max()


def compare_finals(fin, ref, frac=0.995, tol=1e-8):
    same = np.ones(fin.size, dtype=bool)
    for f in ("moves", "cell", "type", "fate", "draws"):
        same &= fin[f] == ref[f]
    assert same.mean() >= frac, f"only {same.mean():.4f} of phonons share the reference's discrete outcome"
    for f in ("time", "pathlen"):
        assert rel_err(fin[f][same], ref[f][same]).max() <= tol, f
    # amplitude = exp(-pi f sum(t/Q)): its relative error is the ABSOLUTE error of an exponent of size |ln amp|
    la, lr = np.log(fin["amp"][same]), np.log(ref["amp"][same])
    assert (np.abs(la - lr) <= tol * (1.0 + np.abs(lr))).all(), "amp"
    scale = max(1.0, np.abs(ref["loc"]).max())
    assert np.abs(fin["loc"][same] - ref["loc"][same]).max() <= tol * scale
    return same


@pytest.mark.parametrize("cfg", CONFIGS)
def test_trace_matches_reference(cfg):
    m, z = load_golden(cfg)
    n, seed = int(z["run_n"]), int(z["run_seed"])
    with engine.Engine(m) as eng:
        fin = eng.trace(n, seed)
        e, c, k = eng.fetch()
    ref = z["run_finals"]
    same = compare_finals(fin, ref)
    e_ref, c_ref = dense_bins(m, z)
    n_off = int((~same).sum())
    # each diverged phonon can move a handful of catches
    assert np.abs(c.astype(np.int64) - c_ref.astype(np.int64)).sum() <= 50 * n_off
    if n_off == 0:
        assert np.array_equal(c, c_ref)
        assert bin_energy_err(e, e_ref) <= 1e-8
        assert np.array_equal(k[:3], z["run_counters"][:3])
    assert int(k[abi.R3D_CNT_PHONONS]) == n
    assert int(k[abi.R3D_CNT_CATCHES]) == int(c.sum())
    assert int(k[:3].sum()) == n


@pytest.mark.parametrize("cfg", ["halfspace", "lopnor"])
def test_run_equals_trace_and_is_additive(cfg):
    """Index-keyed RNG: splitting a range across calls must not change counts (energies up to summation order)."""
    m, z = load_golden(cfg)
    n = 3000
    with engine.Engine(m) as eng:
        eng.run_simulation(n, seed=5)
        e1, c1, k1 = eng.fetch()
        eng.reset()
        eng.run_simulation(1000, seed=5, first_phonon=0)
        eng.run_simulation(1500, seed=5, first_phonon=1000)
        eng.run_simulation(500, seed=5, first_phonon=2500)
        e2, c2, k2 = eng.fetch()
        eng.reset()
        e0, c0, k0 = eng.fetch()
    assert np.array_equal(c1, c2) and np.array_equal(k1[:7], k2[:7])
    assert bin_energy_err(e1, e2) <= 1e-12
    assert c1.sum() > 0 and c0.sum() == 0 and k0.sum() == 0 and e0.sum() == 0


@pytest.mark.parametrize("cfg", CONFIGS)
def test_bigger_run_vs_oracle(cfg):
    """More phonons than the golden fixtures hold, against the oracle (itself pinned to the reference)."""
    m, _ = load_golden(cfg)
    n = {"spherical": 1500}.get(cfg, 20000)
    e_ref, c_ref, k_ref, f_ref = ob.run(m, 100000, n, 99, finals=True, nthreads=4)
    with engine.Engine(m) as eng:
        fin = eng.trace(n, seed=99, first_phonon=100000)
        e, c, k = eng.fetch()
    same = compare_finals(fin, f_ref)
    n_off = int((~same).sum())
    assert np.abs(c.astype(np.int64) - c_ref.astype(np.int64)).sum() <= 50 * n_off
    tot, tot_ref = e[..., 3:].sum(), e_ref[..., 3:].sum()
    assert abs(tot - tot_ref) <= (1e-9 + 1e-3 * n_off) * tot_ref
    # loss / time-out tallies agree up to the diverged phonons
    assert np.abs(k[:3].astype(np.int64) - k_ref[:3].astype(np.int64)).sum() <= 2 * n_off


def test_empty_and_ragged_ranges():
    m, _ = load_golden("halfspace")
    with engine.Engine(m) as eng:
        eng.run_simulation(0)
        eng.run_simulation(1, first_phonon=2**40)
        eng.run_simulation(129)
        t = eng.sync()
        e, c, k = eng.fetch()
        assert int(k[abi.R3D_CNT_PHONONS]) == 130 and int(k[:3].sum()) == 130 and t >= 0
        assert eng.trace(0).size == 0
        assert eng.launch_count >= 2 * 2 + 1      # one propagate launch per non-empty job + one tally reduction per job


def test_no_deflect_and_no_seismometers():
    m, _ = load_golden("halfspace")
    m.no_deflect = 1
    m.seis = np.zeros(0)
    e_ref, c_ref, k_ref, f_ref = ob.run(m, 0, 5000, 3, finals=True)
    with engine.Engine(m) as eng:
        fin = eng.trace(5000, seed=3)
        e, c, k = eng.fetch()
    compare_finals(fin, f_ref)
    assert e.size == 0 and int(k[abi.R3D_CNT_SCATTERS]) == int(k_ref[abi.R3D_CNT_SCATTERS])


def test_device_accumulators_wrap_as_torch():
    import torch
    m, _ = load_golden("halfspace_nearsrc50")
    with engine.Engine(m) as eng:
        eng.run_simulation(20000, seed=11)
        eng.sync()
        e, c, k = eng.fetch()
        de, dc, dk = (torch.as_tensor(v, device="cuda:0") for v in eng.device_accumulators(0))
        assert de.dtype == torch.float64 and dc.dtype == torch.int64
        assert np.array_equal(dc.cpu().numpy().astype(np.uint64).reshape(c.shape), c)
        assert np.array_equal(de.cpu().numpy().reshape(e.shape), e)
        assert int(dk[abi.R3D_CNT_PHONONS]) == 20000


def test_kernel_times_report():
    """The roofline hook: one event-timed launch per job, phase shares that add up, units equal to the counters."""
    m, _ = load_golden("halfspace")
    with engine.Engine(m) as eng:
        eng.set_profiling(True)
        eng.run_simulation(50000, seed=2)
        eng.sync()
        kt = eng.kernel_times()
        e, c, k = eng.fetch()
    assert kt["launches"] == 1 and kt["seconds"] > 0 and kt["iterations"] >= 1 and kt["ctas"] >= 1
    assert abs(kt["phase1_seconds"] + kt["phase2_seconds"] - kt["seconds"]) <= 1e-9 + 1e-6 * kt["seconds"]
    assert kt["events"] == int(k[abi.R3D_CNT_EVENTS]) and kt["catches"] == int(c.sum())
    assert kt["draws"] == int(k[abi.R3D_CNT_SCATTERS]) + 50000


def test_two_devices_in_one_handle():
    """r3d_create(devices[]) replicates the model and r3d_run shards the index range: the result equals the
    single-device one (counts exactly, energies up to summation order)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    m, _ = load_golden("lopnor")
    n = 30001
    with engine.Engine(m, devices=(0,)) as eng:
        eng.run_simulation(n, seed=21)
        e1, c1, k1 = eng.fetch()
    with engine.Engine(m, devices=(0, 1)) as eng:
        eng.run_simulation(n, seed=21)
        e2, c2, k2 = eng.fetch()
    assert np.array_equal(c1, c2) and np.array_equal(k1[:7], k2[:7])
    assert bin_energy_err(e1, e2) <= 1e-12


def test_fetch_into_caller_buffers():
    """r3d_fetch fills the arrays it is given (every element, so stale contents cannot survive) and rejects wrong shapes."""
    m, z = load_golden("halfspace")
    with engine.Engine(m) as eng:
        eng.run_simulation(20000, seed=5)
        eng.sync()
        e, c, k = eng.fetch()
        e2 = np.full(e.shape, 7.0)
        c2 = np.full(c.shape, 9, dtype=np.uint64)
        e3, c3, k3 = eng.fetch(out=(e2, c2))
        assert e3 is e2 and c3 is c2
        assert np.array_equal(e2, e) and np.array_equal(c2, c) and np.array_equal(k3, k)
        with pytest.raises(ValueError):
            eng.fetch(out=(e2[:, :-1], c2))
        with pytest.raises(ValueError):
            eng.fetch(out=(e2.astype(np.float32), c2))


def test_both_builds_of_the_layered_kernel_agree(monkeypatch):
    """The layered models have two builds of the kernel (384 x 168 and 512 x 128 registers, chosen by pilot launches on the
    first large job): same phonons, same draws, same arithmetic - identical counts and counters, energies up to summation order."""
    m, _ = load_golden("lopnor")
    res = []
    for wide in ("0", "1"):
        monkeypatch.setenv("R3D_CYL_WIDE", wide)
        with engine.Engine(m) as eng:
            eng.run_simulation(60000, seed=9)
            res.append(eng.fetch())
    (e0, c0, k0), (e1, c1, k1) = res
    assert int(k0[abi.R3D_CNT_PHONONS]) == 60000 and int(c0.sum()) > 0
    assert np.array_equal(c0, c1) and np.array_equal(k0, k1)
    assert bin_energy_err(e0, e1) <= 1e-12


def test_pilot_launches_are_part_of_the_job(monkeypatch):
    """A job above the pilot threshold (8e6 phonons) starts with one launch of 2e6 phonons per build; together with the rest
    they trace every phonon exactly once: the result equals that of a handle whose build was fixed beforehand."""
    monkeypatch.delenv("R3D_CYL_WIDE", raising=False)
    m, _ = load_golden("halfspace")
    n = 9_000_001
    with engine.Engine(m) as eng:
        launches0 = eng.launch_count
        eng.run_simulation(n, seed=4)
        e0, c0, k0 = eng.fetch()
        assert eng.launch_count - launches0 == 4                 # two pilots, the rest, the tally reduction
        eng.run_simulation(n, seed=4, first_phonon=n)            # the choice is kept: one launch + the tally reduction
        eng.sync()
        assert eng.launch_count - launches0 == 6
    monkeypatch.setenv("R3D_CYL_WIDE", "0")
    with engine.Engine(m) as eng:
        eng.run_simulation(n, seed=4)
        e1, c1, k1 = eng.fetch()
    assert int(k0[abi.R3D_CNT_PHONONS]) == n and int(k0[:3].sum()) == n
    assert np.array_equal(c0, c1) and np.array_equal(k0, k1)
    assert bin_energy_err(e0, e1) <= 1e-12


def test_job_larger_than_one_launch():
    """A launch holds at most 2^30 phonons (32-bit index in the slot, 32-bit per-thread tallies); a larger job is cut into
    launches and still traces every phonon once."""
    m, _ = load_golden("halfspace")
    n = (1 << 30) + 12345
    with engine.Engine(m) as eng:
        eng.run_simulation(n, seed=2)
        t = eng.sync()
        e, c, k = eng.fetch()
    assert int(k[abi.R3D_CNT_PHONONS]) == n and int(k[:3].sum()) == n
    assert t < 5.0


def test_large_tables_in_device_memory():
    """r3d_create takes the five large tables from device memory (what a rank gets from the NCCL broadcast of
    distributed.broadcast_model_tables): same result as from host memory."""
    import torch
    from radiative3d_b200 import distributed
    m, _ = load_golden("crustpinch")
    with engine.Engine(m) as eng:
        eng.run_simulation(20000, seed=6)
        e0, c0, k0 = eng.fetch()
    tables = distributed.broadcast_model_tables(m, torch.device("cuda", 0))      # (no process group: upload only)
    torch.cuda.synchronize()
    with engine.Engine(m, device_tables=tables) as eng:
        eng.run_simulation(20000, seed=6)
        e1, c1, k1 = eng.fetch()
    assert np.array_equal(c0, c1) and np.array_equal(k0, k1)
    assert bin_energy_err(e0, e1) <= 1e-12            # (the order of the bin reductions differs from run to run)
    with pytest.raises(ValueError):
        engine.Engine(m, device_tables={"cell_params": 1234})
