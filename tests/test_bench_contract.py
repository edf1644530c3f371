"""bench.py keeps the measurement contract: the reference arm is run here for real (CPU, a small sample), the GPU arm on the
box (-m gpu, a small batch), and the bench lines recorded under profiles/ for this round carry the same keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_main")


def check_gpu_line(d, workload):
    for k in BASE:
        assert k in d, k
    assert d["metric"] == "phonons traced/sec" and d["unit"] == "phonons/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None and d["warmup"] >= 3
    assert "workload" in d["config"] and workload in d["config"]["workload"]
    assert d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "what"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        assert c["kind"] in ("reference", "port")


def check_reference_line(d):
    assert d["impl"] == "reference" and d["metric"] == "phonons traced/sec" and d["unit"] == "phonons/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["higher_is_better"] is True and "workload" in d["config"]


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="oracle/_ref/r3d_ref_main not built")
@pytest.mark.parametrize("workload", ["halfspace_nearsrc50", "lopnor"])
def test_reference_arm_runs(workload):
    """bench.py --impl reference, for real: two processes of the unmodified reference binary, a small sample, TOA degree 3."""
    env = dict(os.environ, R3D_BENCH_TOA_DEGREE="3", R3D_BENCH_REF_CORES="2", R3D_BENCH_REF_PER_PROC="2000" if workload == "lopnor" else "20000")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", workload], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    check_reference_line(d)
    assert d["cpu_baseline"]["cores"] == 2 and workload in d["config"]["workload"]
    assert "__BEGINNING_SIMULATION__" in d["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_stay_silent():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True, text=True,
                       timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_roofline_helpers():
    import bench
    assert bench.algorithmic_bytes_per_draw(5242880) == 23 * 8 + 16       # 200 B per draw at TOA degree 9
    assert bench.algorithmic_bytes_per_draw(320) == 9 * 8 + 16
    assert set(bench.WORKLOADS) == {"halfspace", "halfspace_nearsrc50", "crustpinch", "lopnor", "spherical"}
    assert bench.WORKLOAD == "halfspace_nearsrc50" and bench.WORKLOADS[bench.WORKLOAD]["per_gpu"] == 125_000_000
    kc = os.path.join(ROOT, "profiles", "kernel_counters.json")
    if os.path.exists(kc):
        r = bench.issue_roofline("halfspace_nearsrc50", 8.5e9, 1900.0)
        assert r["bound"] == "issue" and 0.0 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["halfspace_nearsrc50", "spherical"])
def test_gpu_arm_runs(workload):
    """bench.py on the box with a small batch: one JSON line on stdout with every key of the contract."""
    n = "4000000" if workload == "halfspace_nearsrc50" else "50000"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--per-gpu", n, "--e2e-steps", "1",
                        "--no-cpu-baseline", "--workload", workload], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, R3D_BENCH_TOA_DEGREE="6"))
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    check_gpu_line(json.loads(lines[0]), workload)


def recorded(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.readline())


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("r2_bench_") and f.endswith(".json")))
def test_recorded_lines_of_this_round(name):
    d = recorded(name)
    if d.get("impl") == "reference":
        check_reference_line(d)
    else:
        wl = d["config"]["workload"].split(" ")[0]
        check_gpu_line(d, wl)
        assert "roofline_issue" in d
