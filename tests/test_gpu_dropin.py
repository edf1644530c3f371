"""The drop-in program (integration/_build/r3d_gpu_main = the reference's own main(), model build and output writers with
Model::RunSimulation() replaced by the C-ABI calls) against the same run made directly through the ABI.

What this pins: the reference-side flattener (integration/r3d_flatten.hpp), the write-back of bins and counters into the
reference's Seismometer / DataReporter objects, and that the reference's writers then produce their usual files
(seis_NNN.octv, dataout.cpp:284-406; loss summary on stdout, dataout.cpp:630-634) from GPU results."""
import os
import re

import numpy as np
import pytest

from radiative3d_b200 import engine, reference_host

pytestmark = pytest.mark.gpu

sys_path_golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_octv(path):
    import sys
    sys.path.insert(0, sys_path_golden)
    from make_stat_golden import read_octv as r
    return r(path)


@pytest.mark.parametrize("cfg,deg,n", [("halfspace", 4, 300000), ("lopnor", 3, 20000), ("spherical", 3, 2000)])
def test_reference_program_with_gpu_loop(cfg, deg, n, tmp_path):
    if not os.path.exists(reference_host.GPU_MAIN):
        pytest.skip("integration/_build/r3d_gpu_main was not built (needs the reference checkout at build time)")
    seed = 4242
    p = reference_host.run(cfg, n, deg, str(tmp_path), seed=seed)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lost = int(re.search(r"Loss surfaces:\s+(\d+)", p.stdout).group(1))
    tmo = int(re.search(r"Timeout:\s+(\d+)", p.stdout).group(1))
    inv = int(re.search(r"Invalidity:\s+(\d+)", p.stdout).group(1))
    assert "@@ __SIMULATION_COMPLETE__" in p.stdout
    files = sorted(f for f in os.listdir(tmp_path) if re.fullmatch(r"seis_\d+\.octv", f))

    m = reference_host.build_model(cfg, deg)
    assert len(files) == m.n_seis
    with engine.Engine(m) as eng:
        # the program traces the range in ten slices (progress lines); slicing does not change the result
        eng.run_simulation(n, seed=seed)
        e, c, k = eng.fetch()
    assert (lost, tmo, inv) == tuple(int(x) for x in k[:3]) and lost + tmo + inv == n
    for i in (0, len(files) // 2, len(files) - 1):
        o = read_octv(os.path.join(tmp_path, files[i]))
        assert np.array_equal(o["CountPS"].astype(np.uint64), c[i])
        # the writers print 6 significant digits
        assert np.allclose(o["TracePS"], e[i][:, 3:5], rtol=2e-5, atol=0)
        assert np.allclose(o["TraceXYZ"], e[i][:, 0:3], rtol=2e-5, atol=1e-300)
    assert os.path.exists(os.path.join(tmp_path, "seis_traces_asc.dat"))


def test_checkpoint_and_resume(tmp_path):
    """R3D_GPU_CHECKPOINT: a run stopped after 4 tenths and resumed equals the uninterrupted run (phonon i always uses
    draw stream (seed, i)): counts and loss counters exactly, energies up to summation order."""
    if not os.path.exists(reference_host.GPU_MAIN):
        pytest.skip("integration/_build/r3d_gpu_main was not built (needs the reference checkout at build time)")
    import subprocess
    from radiative3d_b200 import workloads
    cfg, deg, n, seed = "halfspace", 4, 400000, 99

    def run(outdir, **extra):
        os.makedirs(outdir, exist_ok=True)
        env = dict(os.environ, R3D_GPU_SEED=str(seed), **extra)
        return subprocess.run([reference_host.GPU_MAIN] + workloads.cmdline(cfg, n, deg, str(outdir)), cwd=str(outdir), env=env,
                              capture_output=True, text=True)

    whole = run(tmp_path / "whole")
    assert whole.returncode == 0
    ck = str(tmp_path / "run.ckpt")
    part = run(tmp_path / "resumed", R3D_GPU_CHECKPOINT=ck, R3D_GPU_STOP_AFTER="4")
    assert part.returncode == 3 and os.path.exists(ck)
    rest = run(tmp_path / "resumed", R3D_GPU_CHECKPOINT=ck)
    assert rest.returncode == 0 and "resuming from checkpoint" in rest.stderr
    summary = lambda p: re.findall(r"(Loss surfaces|Timeout|Invalidity):\s+(\d+)", p.stdout)
    assert summary(whole) == summary(rest)
    for i in (0, 70, 143):
        a = read_octv(os.path.join(tmp_path / "whole", f"seis_{i:03d}.octv"))
        b = read_octv(os.path.join(tmp_path / "resumed", f"seis_{i:03d}.octv"))
        assert np.array_equal(a["CountPS"], b["CountPS"])
        assert np.allclose(a["TracePS"], b["TracePS"], rtol=1e-5, atol=0)
