"""N > 1 host logic on CPU: two gloo ranks shard a phonon range, each traces its share (with the oracle standing
in for a device), the bins are all-reduced, and the result must equal the single-process run."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, load_golden
from radiative3d_b200 import distributed


def test_shard_range_covers_exactly():
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            parts = [distributed.shard_range(5, n, r, world) for r in range(world)]
            assert parts[0][0] == 5 and sum(p[1] for p in parts) == n
            for a, b in zip(parts, parts[1:]):
                assert a[0] + a[1] == b[0]
            assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import oracle_binding as ob
    from conftest import load_golden
    from radiative3d_b200 import distributed
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    m, _ = load_golden("lopnor")
    first, n = distributed.shard_range(1000, 901, rank, world)
    e, c, k, _ = ob.run(m, first, n, 77)
    if rank == 1:
        k[7] |= np.uint64(1 << 4)                 # pretend rank 1 saw a STUCK phonon: the OR must survive
    # the same reduction on a handle's two accumulator blocks (include/r3d_gpu.h: r3d_device_accumulator_blocks), to rank 0
    import torch
    from radiative3d_b200 import abi
    lanes = np.array([(int(k[7]) >> b) & 1 for b in range(abi.R3D_NDIAG_LANES)], dtype=np.int64)
    fblk = torch.from_numpy(e.reshape(-1).copy())
    iblk = torch.from_numpy(np.concatenate([c.reshape(-1).view(np.int64), k.view(np.int64), lanes]))
    distributed.all_reduce_blocks(fblk, iblk, c.size, dst=0)
    distributed.all_reduce_numpy(e, c, k)
    np.savez({out!r} + f".{{rank}}.npz", e=e, c=c, k=k, fblk=fblk.numpy(), iblk=iblk.numpy())
    dist.destroy_process_group()
""")


def test_two_ranks_match_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    out = str(tmp_path / "res")
    script.write_text(WORKER.format(root=ROOT, out=out))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    import oracle_binding as ob
    m, _ = load_golden("lopnor")
    e, c, k, _ = ob.run(m, 1000, 901, 77)
    for r in range(2):
        z = np.load(out + f".{r}.npz")
        assert np.array_equal(z["c"], c)
        assert np.array_equal(z["k"][:7], k[:7])
        assert int(z["k"][7]) == 1 << 4
        assert np.abs(z["e"] - e).max() <= 1e-12 * max(1.0, np.abs(e).max())
    assert c.sum() > 0
    # two collectives on the blocks, reduced to rank 0: same numbers, diagnostic word rebuilt from its lanes
    z = np.load(out + ".0.npz")
    assert np.array_equal(z["iblk"][:c.size].view(np.uint64), c.reshape(-1))
    assert np.array_equal(z["iblk"][c.size:c.size + 7].view(np.uint64), k[:7])
    assert int(z["iblk"][c.size + 7]) == 1 << 4
    assert np.abs(z["fblk"] - e.reshape(-1)).max() <= 1e-12 * max(1.0, np.abs(e).max())
