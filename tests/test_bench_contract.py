"""The recorded bench lines (profiles/) carry every key the measurement contract names (bench.py's docstring; no GPU needed)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.readline())


@pytest.mark.parametrize("name", ["r1_bench.json", "r1_bench_2gpu.json", "r1_bench_4gpu.json", "r1_bench_8gpu.json"])
def test_bench_line(name):
    d = load(name)
    for k in BASE:
        assert k in d, k
    assert d["metric"] == "phonons traced/sec" and d["unit"] == "phonons/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None and d["warmup"] >= 3
    assert "workload" in d["config"] and "halfspace_nearsrc50" in d["config"]["workload"]
    assert d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        assert c["kind"] in ("reference", "port")


def test_reference_line():
    d = load("r1_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "phonons traced/sec" and d["unit"] == "phonons/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
