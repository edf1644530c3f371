"""The reference's own host program as the model builder of the GPU path.

Model construction (command line, user_*_inc.cpp plugins, grid, cells, scatterer tables, source,
seismometers) stays the reference's C++ (BASELINE.json north_star).  integration/_build/r3d_gpu_main is
that program, compiled from the unmodified reference sources, with Model::RunSimulation() replaced by the
C-ABI calls (integration/r3d_run_simulation_gpu.cpp).  Here it is used in two ways:
  * build_model(): run it with R3D_GPU_DUMP_ONLY so that it stops after writing the flattened model,
  * run():         run a whole simulation through it (reference CLI in, reference output files out).
Nothing here touches oracle/.
"""
import os
import subprocess
import tempfile

from . import workloads
from .model import FlatModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_MAIN = os.path.join(ROOT, "integration", "_build", "r3d_gpu_main")


def _need_binary():
    if not os.path.exists(GPU_MAIN):
        raise FileNotFoundError(
            f"{GPU_MAIN} is missing. It is built by `make -C integration` (called from __graft_entry__.build()) "
            "where the reference checkout exists, and travels to the GPU box as a prebuilt file.")


def build_model(config, toa_degree, extra_args=()):
    """Flattened model of a named workload (radiative3d_b200.workloads.CONFIGS) at the given TOA degree."""
    _need_binary()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "model.r3dmodel")
        env = dict(os.environ, R3D_GPU_DUMP_MODEL=path, R3D_GPU_DUMP_ONLY="1")
        p = subprocess.run([GPU_MAIN] + workloads.cmdline(config, 10, toa_degree, tmp, extra_args), cwd=tmp, env=env,
                           capture_output=True, text=True)
        if p.returncode != 0 or not os.path.exists(path):
            raise RuntimeError(f"model build failed (rc {p.returncode}):\n{p.stdout[-1500:]}\n{p.stderr[-1500:]}")
        return FlatModel.load(path)


def run(config, n_phonons, toa_degree, outdir, seed=None, devices=(0,), extra_args=()):
    """A whole run through the reference CLI with the GPU path inside; returns the CompletedProcess.
    Output files (seis_NNN.octv, seis_traces_asc.dat, stdout summary) are the reference's own formats."""
    _need_binary()
    os.makedirs(outdir, exist_ok=True)
    env = dict(os.environ, R3D_GPU_DEVICES=",".join(str(d) for d in devices), R3D_GPU_NUM_PHONONS=str(int(n_phonons)))
    if seed is not None:
        env["R3D_GPU_SEED"] = str(int(seed))
    return subprocess.run([GPU_MAIN] + workloads.cmdline(config, min(int(n_phonons), 2**31 - 1), toa_degree, outdir, extra_args),
                          cwd=outdir, env=env, capture_output=True, text=True)
