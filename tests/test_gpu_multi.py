"""N > 1 on real GPUs (needs two; skipped on a one-GPU box -- profiles/r2_multi_gpu_tests.log holds a 2-GPU run).

  * in-process: r3d_create(devices[2]) replicates the model by peer copies, r3d_run shards the index range, r3d_fetch sums
    the devices' bins with a kernel over peer memory; with R3D_PEER_COMBINE=0 the host sum is used instead - all three
    must agree with the one-device run (counts and counters exactly, energies to 1e-12: the reference's combine.m:26-33);
  * multi-process: two ranks under torch.distributed.run with NCCL: the model's large tables are uploaded by rank 0 and
    broadcast over NVLink, each rank traces its share, the accumulator blocks are reduced with two collectives, and rank
    0's r3d_fetch must return the one-device result (the diagnostic word OR-ed, not summed).
"""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT, load_golden
from radiative3d_b200 import abi, engine

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _energy_err(a, b):
    scale = max(1.0, float(np.abs(a).max()))
    return float(np.abs(a - b).max()) / scale


@pytest.mark.parametrize("cfg,n", [("lopnor", 30001), ("halfspace", 200003), ("spherical", 3001)])
def test_two_devices_in_one_handle_peer_and_host_sum(cfg, n):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    m, _ = load_golden(cfg)
    with engine.Engine(m, devices=(0,)) as eng:
        eng.run_simulation(n, seed=21)
        e1, c1, k1 = eng.fetch()
    with engine.Engine(m, devices=(0, 1)) as eng:
        eng.run_simulation(n, seed=21)
        e2, c2, k2 = eng.fetch()
        e2b, c2b, k2b = eng.fetch()                # fetching twice must not change anything (sums go to a staging buffer)
    os.environ["R3D_PEER_COMBINE"] = "0"
    os.environ["R3D_PEER_REPLICATE"] = "0"
    try:
        with engine.Engine(m, devices=(1, 0)) as eng:
            eng.run_simulation(n, seed=21)
            e3, c3, k3 = eng.fetch()
    finally:
        del os.environ["R3D_PEER_COMBINE"], os.environ["R3D_PEER_REPLICATE"]
    assert int(k1[abi.R3D_CNT_PHONONS]) == n and int(c1.sum()) > 0
    for (e, c, k) in ((e2, c2, k2), (e2b, c2b, k2b), (e3, c3, k3)):
        assert np.array_equal(c1, c) and np.array_equal(k1, k)
        assert _energy_err(e1, e) <= 1e-12


def test_two_devices_diag_is_ored():
    """An INVALID phonon on either device: counters sum, the diagnostic word is OR-ed (dataout.cpp:611-617)."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    m, z = load_golden("stuck", prefix="inv")
    n, seed = int(z["run_n"]), int(z["run_seed"])
    with engine.Engine(m, devices=(0,)) as eng:
        eng.run_simulation(n, seed=seed)
        _, _, k1 = eng.fetch()
    with engine.Engine(m, devices=(0, 1)) as eng:
        eng.run_simulation(n, seed=seed)
        _, _, k2 = eng.fetch()
    assert np.array_equal(k1, k2)
    assert int(k1[abi.R3D_CNT_INVALID]) > 0 and int(k1[abi.R3D_CNT_DIAG]) == 1 << abi.R3D_INV_STUCK


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from conftest import load_golden
    from radiative3d_b200 import abi, distributed, engine
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {{}}
    for cfg, n, seed in (("lopnor", 30001, 21), ("halfspace", 200003, 5), ("inv_stuck", 0, 0)):
        if cfg.startswith("inv_"):
            m, z = load_golden(cfg[4:], prefix="inv")
            n, seed = int(z["run_n"]), int(z["run_seed"])
        else:
            m, z = load_golden(cfg)
        if rank != 0:                                  # the large tables exist on rank 0's host only
            for name in distributed.BIG_TABLES:
                getattr(m, name)[...] = 0.0
        tables = distributed.broadcast_model_tables(m, torch.device("cuda", local))
        torch.cuda.synchronize()
        with engine.Engine(m, devices=(local,), device_tables=tables) as eng:
            first, cnt = distributed.shard_range(100, n, rank, world)
            eng.run_simulation(cnt, seed=seed, first_phonon=first)
            eng.sync()
            f, i, at = eng.device_accumulator_blocks(0)
            tf, ti = torch.as_tensor(f, device=f"cuda:{{local}}"), torch.as_tensor(i, device=f"cuda:{{local}}")
            distributed.all_reduce_blocks(tf, ti, at, dst=0)
            torch.cuda.synchronize()
            if rank == 0:
                e, c, k = eng.fetch()
                out[cfg + "_e"], out[cfg + "_c"], out[cfg + "_k"] = e, c, k
        del tables
    if rank == 0:
        np.savez({out!r}, **out)
    dist.destroy_process_group()
""")


def test_two_ranks_nccl_match_single_gpu(tmp_path):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    out = str(tmp_path / "res.npz")
    script.write_text(WORKER.format(root=ROOT, out=out))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    z = np.load(out)
    for cfg, n, seed in (("lopnor", 30001, 21), ("halfspace", 200003, 5), ("inv_stuck", 0, 0)):
        m, zz = load_golden(cfg[4:], prefix="inv") if cfg.startswith("inv_") else load_golden(cfg)
        if cfg.startswith("inv_"):
            n, seed = int(zz["run_n"]), int(zz["run_seed"])
        with engine.Engine(m, devices=(0,)) as eng:
            eng.run_simulation(n, seed=seed, first_phonon=100)
            e, c, k = eng.fetch()
        assert int(k[abi.R3D_CNT_PHONONS]) == n
        assert np.array_equal(z[cfg + "_c"], c), cfg
        assert np.array_equal(z[cfg + "_k"], k), cfg           # includes the diagnostic word: OR, not sum
        assert _energy_err(e, z[cfg + "_e"]) <= 1e-12, cfg
        if cfg == "inv_stuck":
            assert int(k[abi.R3D_CNT_INVALID]) > 0 and int(k[abi.R3D_CNT_DIAG]) == 1 << 4
