#!/usr/bin/env python
"""profiles/kernel_counters.json from `ncu --set full` captures of the propagate kernel (run here, no GPU needed).

    python scripts/ncu_counters.py <tag> <workload>=<file.ncu-rep>:<log of the same profile_target.py run> ...

Each capture is ONE launch of propagate_kernel (scripts/r2_session_*.sh: the second launch of profile_target.py); the log's
last line gives the phonons of that launch and the loop events per phonon.  bench.py reads the result for `roofline.traffic`
(DRAM bytes per phonon) and `roofline_issue` (warp instructions per loop event)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
path = os.environ.get("R3D_COUNTERS_OUT") or os.path.join(ROOT, "profiles", "kernel_counters.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for spec in sys.argv[2:]:
    wl, rest = spec.split("=", 1)
    rep, log = rest.split(":", 1)
    line = [ln for ln in open(log) if "phonons/s" in ln][-1]
    n = int(float(re.search(r"n=(\d+)", line).group(1)))
    ev = float(re.search(r"\[([\d.]+) events", line).group(1))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))

    def val(k):
        v = float(d[k].replace(",", ""))
        unit = u.get(k, "")
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)

    inst = val("smsp__inst_executed.sum")
    lanes = val("smsp__thread_inst_executed_per_inst_executed.ratio")
    events = n * ev
    out[wl] = {
        "warp_inst_per_event": inst / events, "thread_inst_per_event": inst * lanes / events, "lanes_per_inst": lanes,
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_per_sm": val("sm__warps_active.avg.pct_of_peak_sustained_active") * 64 / 100,
        "dram_bytes_per_phonon": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / n,
        "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
        "red_sectors_per_phonon": val("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum") / n if "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum" in d else None,
        "registers": int(val("launch__registers_per_thread")), "block_size": int(val("launch__block_size")),
        "kernel_ms_under_ncu": val("gpu__time_duration.sum") * (1e-6 if u.get("gpu__time_duration.sum") == "ns" else 1e-3 if u.get("gpu__time_duration.sum") == "us" else 1.0),
        "phonons_in_launch": n, "events_per_phonon": ev,
        "source": f"ncu --set full --clock-control none, one launch of {n} phonons at TOA degree 9 ({tag}; summary in profiles/{tag}_{wl}_kernel.md)",
    }
    print(wl, json.dumps(out[wl], indent=1))
json.dump(out, open(path, "w"), indent=1)
