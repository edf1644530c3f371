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
    """--gpu-checkpoint: a run stopped after 4 tenths and resumed equals the uninterrupted run (phonon i always uses draw
    stream (seed, i)): counts and loss counters exactly, energies up to summation order.  The resumed run is started
    WITHOUT --seed and takes it from the checkpoint; a checkpoint of another run is refused, not overwritten."""
    if not os.path.exists(reference_host.GPU_MAIN):
        pytest.skip("integration/_build/r3d_gpu_main was not built (needs the reference checkout at build time)")
    import subprocess
    from radiative3d_b200 import workloads
    cfg, deg, n, seed = "halfspace", 4, 400000, 99

    def run(outdir, *opts, count=n, **extra):
        os.makedirs(outdir, exist_ok=True)
        return subprocess.run([reference_host.GPU_MAIN] + workloads.cmdline(cfg, count, deg, str(outdir)) + list(opts), cwd=str(outdir),
                              env=dict(os.environ, **extra), capture_output=True, text=True)

    whole = run(tmp_path / "whole", f"--seed={seed}")
    assert whole.returncode == 0
    ck = str(tmp_path / "run.ckpt")
    part = run(tmp_path / "resumed", f"--seed={seed}", f"--gpu-checkpoint={ck}", R3D_GPU_STOP_AFTER="4")
    assert part.returncode == 3 and os.path.exists(ck)
    before = open(ck, "rb").read()
    # another run must not take the file over: different phonon count, and different seed
    other = run(tmp_path / "other", f"--seed={seed}", f"--gpu-checkpoint={ck}", count=n + 1)
    assert other.returncode == 1 and "belongs to another run" in other.stdout + other.stderr
    other = run(tmp_path / "other", f"--seed={seed + 1}", f"--gpu-checkpoint={ck}")
    assert other.returncode == 1 and "belongs to another run" in other.stdout + other.stderr
    assert open(ck, "rb").read() == before
    rest = run(tmp_path / "resumed", f"--gpu-checkpoint={ck}")                # no --seed: taken from the checkpoint
    assert rest.returncode == 0 and "resuming from checkpoint" in rest.stderr and f"seed {seed}" in rest.stderr
    summary = lambda p: re.findall(r"(Loss surfaces|Timeout|Invalidity):\s+(\d+)", p.stdout)
    assert summary(whole) == summary(rest)
    for i in (0, 70, 143):
        a = read_octv(os.path.join(tmp_path / "whole", f"seis_{i:03d}.octv"))
        b = read_octv(os.path.join(tmp_path / "resumed", f"seis_{i:03d}.octv"))
        assert np.array_equal(a["CountPS"], b["CountPS"])
        assert np.allclose(a["TracePS"], b["TracePS"], rtol=1e-5, atol=0)


def test_command_line_options_of_the_gpu_path(tmp_path):
    """--num-phonons with the reference's K suffix read as a 64-bit count, --seed, --gpu-devices (integration/r3d_cli.hpp):
    the run equals the same range through the ABI, and out_mparams.octv carries the count (model.cpp:148-150)."""
    if not os.path.exists(reference_host.GPU_MAIN):
        pytest.skip("integration/_build/r3d_gpu_main was not built (needs the reference checkout at build time)")
    import subprocess
    from radiative3d_b200 import workloads
    cfg, deg, seed = "halfspace_nearsrc50", 3, 31337
    args = [a for a in workloads.cmdline(cfg, 10, deg, str(tmp_path)) if not a.startswith("--num-phonons")]
    args += ["-N", "250K", f"--seed={seed}", "--gpu-devices", "0", "--mparams-outfile=out_mparams.octv"]
    p = subprocess.run([reference_host.GPU_MAIN] + args, cwd=str(tmp_path), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "250000 phonons on 1 device(s)" in p.stderr and f"seed {seed}" in p.stderr
    txt = open(os.path.join(tmp_path, "out_mparams.octv")).read()
    assert re.search(r"# name: NumPhonons \n# type: scalar \n250000 ", txt)
    m = reference_host.build_model(cfg, deg)
    with engine.Engine(m) as eng:
        eng.run_simulation(250000, seed=seed)
        e, c, k = eng.fetch()
    got = tuple(int(re.search(rf"{name}:\s+(\d+)", p.stdout).group(1)) for name in ("Loss surfaces", "Timeout", "Invalidity"))
    assert got == tuple(int(x) for x in k[:3])
    o = read_octv(os.path.join(tmp_path, "seis_010.octv"))
    assert np.array_equal(o["CountPS"].astype(np.uint64), c[10])
