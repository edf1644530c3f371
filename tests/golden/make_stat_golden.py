#!/usr/bin/env python
"""Generate tests/golden/stat_<cfg>.npz: whole runs of the UNMODIFIED reference program (oracle/_ref/r3d_ref_main) with its
own rand() stream, for the statistical parity test (tests/test_gpu_statistical.py; SURVEY 8d "statistical acceptance").

    python tests/golden/make_stat_golden.py [cfg ...]

Each workload is run as P independent processes (the reference's own way of scaling, scripts/do-parallel.sh) started more than
a second apart so that srand(time(NULL)) (model.cpp:235) seeds them differently, at the take-off-angle degree of the matching
golden_<cfg>.npz model.  Stored: window-summed seismometer counts and energies per process (the per-process scatter is what the
test uses as the Monte-Carlo error of an energy), and the loss counters the program prints.
"""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from radiative3d_b200 import workloads  # noqa: E402
from make_golden import PLAN as MODEL_PLAN  # noqa: E402

REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_main")
WINDOWS = 10          # time windows per seismometer
# fixture -> (workload, take-off-angle degree (None: that of golden_<workload>.npz), processes, phonons per process)
PLAN = {
    "halfspace": ("halfspace", None, 16, 1_000_000),
    "halfspace_nearsrc50": ("halfspace_nearsrc50", None, 16, 1_000_000),
    "crustpinch": ("crustpinch", None, 16, 20_000),
    "lopnor": ("lopnor", None, 16, 20_000),
    "spherical": ("spherical", None, 16, 1_200),
    # the benchmarked configuration itself (bench.py): the scripted degree 9, 5 242 880 take-off angles; the GPU test builds
    # the same model on the box with the reference's host code (integration/_build/r3d_gpu_main)
    "halfspace_nearsrc50_deg9": ("halfspace_nearsrc50", 9, 16, 1_000_000),
}


def read_octv(path):
    """The matrices of one seis_NNN.octv (dataout.cpp:284-406) as {name: ndarray}."""
    out, name, rows, cols, buf = {}, None, None, None, []
    for line in open(path):
        s = line.strip()
        if s.startswith("# name:"):
            name, rows, cols, buf = s.split(":", 1)[1].strip(), None, None, []
        elif s.startswith("# rows:"):
            rows = int(s.split(":")[1])
        elif s.startswith("# columns:"):
            cols = int(s.split(":")[1])
        elif s and not s.startswith("#") and rows is not None and cols is not None and name not in out:
            buf.extend(float(x) for x in s.split())
            if len(buf) >= rows * cols:
                out[name] = np.array(buf[:rows * cols]).reshape(rows, cols)
    return out


def run_batch(cfg, deg, procs, n_each, tmp):
    ps = []
    for i in range(procs):
        d = os.path.join(tmp, f"p{i}")
        os.makedirs(d)
        ps.append((d, subprocess.Popen([REF_MAIN] + workloads.cmdline(cfg, n_each, deg, d), cwd=d, stdout=subprocess.PIPE,
                                       stderr=subprocess.DEVNULL, text=True)))
        time.sleep(1.1)                      # a different time(NULL) for the next process
    res = []
    for d, p in ps:
        out = p.communicate()[0]
        if p.returncode != 0:
            raise RuntimeError(f"reference failed in {d}")
        lost = int(re.search(r"Loss surfaces:\s+(\d+)", out).group(1))
        tmo = int(re.search(r"Timeout:\s+(\d+)", out).group(1))
        inv = int(re.search(r"Invalidity:\s+(\d+)", out).group(1))
        files = sorted(f for f in os.listdir(d) if re.fullmatch(r"seis_\d+\.octv", f))
        cnt, en = [], []
        for f in files:
            m = read_octv(os.path.join(d, f))
            cnt.append(m["CountPS"])
            en.append(m["TracePS"])
        res.append((np.array(cnt), np.array(en), (lost, tmo, inv)))
    return res


def windows(a, w=WINDOWS):
    """[n_seis, n_bins, k] -> [n_seis, w, k] sums over w equal groups of bins (the last takes the remainder)."""
    n = a.shape[1]
    edges = [n * i // w for i in range(w + 1)]
    return np.stack([a[:, edges[i]:edges[i + 1]].sum(axis=1) for i in range(w)], axis=1)


def main():
    if not os.path.exists(REF_MAIN):
        sys.exit("oracle/_ref/r3d_ref_main is missing: run `make -C oracle ref` where /root/reference exists")
    for name in (sys.argv[1:] or list(PLAN)):
        cfg, deg, procs, n_each = PLAN[name]
        if deg is None:
            deg = MODEL_PLAN[cfg][0]
        t = time.time()
        with tempfile.TemporaryDirectory() as tmp:
            res = run_batch(cfg, deg, procs, n_each, tmp)
        counts = np.stack([windows(c) for c, _, _ in res]).astype(np.int64)          # [P, n_seis, W, 2]
        energy = np.stack([windows(e) for _, e, _ in res])                           # [P, n_seis, W, 2]
        counters = np.array([k for _, _, k in res], dtype=np.int64)                   # [P, 3] lost, timeout, invalid
        path = os.path.join(HERE, f"stat_{name}.npz")
        np.savez_compressed(path, counts=counts, energy=energy.astype(np.float32), counters=counters,
                            n_each=np.int64(n_each), toa_degree=np.int64(deg), windows=np.int64(WINDOWS))
        print(f"stat_{name}.npz: {cfg} at TOA degree {deg}, {procs} x {n_each} phonons, {int(counts.sum())} catches, counters {counters.sum(axis=0)}, "
              f"{os.path.getsize(path) / 1e3:.0f} kB, {time.time() - t:.0f} s")


if __name__ == "__main__":
    main()
