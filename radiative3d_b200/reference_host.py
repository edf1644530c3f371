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


def _table_env(gpu_tables):
    """Scatterer tables of the model build: on the GPU (integration/r3d_scatterer_gpu.cpp) wherever there is one.  Only the
    model BUILD has a CPU form (the reference's own GSATO loop, for building models on a machine without a GPU, e.g. to
    inspect them); the propagate path has none."""
    if gpu_tables is None:
        gpu_tables = os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")
    return {} if gpu_tables else {"R3D_GPU_SCATTERERS": "0"}


def build_model(config, toa_degree, extra_args=(), gpu_tables=None):
    """Flattened model of a named workload (radiative3d_b200.workloads.CONFIGS) at the given TOA degree."""
    _need_binary()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "model.r3dmodel")
        env = dict(os.environ, R3D_GPU_DUMP_MODEL=path, R3D_GPU_DUMP_ONLY="1", **_table_env(gpu_tables))
        p = subprocess.run([GPU_MAIN] + workloads.cmdline(config, 10, toa_degree, tmp, extra_args), cwd=tmp, env=env,
                           capture_output=True, text=True)
        if p.returncode != 0 or not os.path.exists(path):
            raise RuntimeError(f"model build failed (rc {p.returncode}):\n{p.stdout[-1500:]}\n{p.stderr[-1500:]}")
        return FlatModel.load(path)


def run(config, n_phonons, toa_degree, outdir, seed=None, devices=(0,), extra_args=(), env=None):
    """A whole run through the reference CLI with the GPU path inside; returns the CompletedProcess.
    Output files (seis_NNN.octv, seis_traces_asc.dat, stdout summary) are the reference's own formats.
    The phonon count (64-bit), the seed and the devices go through the options of integration/r3d_cli.hpp."""
    _need_binary()
    os.makedirs(outdir, exist_ok=True)
    args = workloads.cmdline(config, int(n_phonons), toa_degree, outdir, extra_args)
    args.append("--gpu-devices=" + ",".join(str(d) for d in devices))
    if seed is not None:
        args.append(f"--seed={int(seed)}")
    return subprocess.run([GPU_MAIN] + args, cwd=outdir, env=dict(os.environ, **(env or {})), capture_output=True, text=True)
