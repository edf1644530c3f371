#!/usr/bin/env python
"""bench.py -- phonons traced per second on the BASELINE.json workload (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the propagate path over one batch of phonons: PER_GPU phonons on every GPU (weak scaling;
1.25e8 = the 1e9 phonons of BASELINE.json configs[1] over the 8 GPUs it is quoted on), on the Halfspace
near-source model (do-halfspace-nearsrc50.sh) at the scripted take-off-angle degree 9.  The model is built by the
reference's own host code (integration/_build/r3d_gpu_main, see radiative3d_b200/reference_host.py).

  value    device-timed (CUDA events on the launching stream, max over ranks): model resident in HBM, K steps plus
           the single end-of-run NCCL all-reduce of the bins.
  e2e      the same batch through the C ABI with HOST buffers: r3d_create (H2D of the pinned model tables) +
           r3d_run + r3d_fetch (D2H of bins and counters) + r3d_destroy, wall clock.
  roofline the propagate kernel (one launch per step and GPU) timed with CUDA events on its stream over one more step.
  cpu_baseline  the unmodified reference binary (oracle/_ref/r3d_ref_main) on one host core, bounded sample.

--impl reference times the reference's own CPU implementation with all host threads it can use (independent
processes, the reference's own way of scaling: scripts/do-parallel.sh), on a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "halfspace_nearsrc50"
TOA_DEGREE = 9
PER_GPU = 125_000_000
SEED = 20261018
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "r3d_ref_main")
# bytes per table draw = ceil(log2 nTOA)*8 + 16, per catch = 96
BYTES_PER_CATCH = 96


def algorithmic_bytes_per_draw(n_toa):
    return (max(1, (n_toa - 1).bit_length())) * 8 + 16


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, throttle reasons and power sampled every 10 ms through NVML while the timed region runs (nvidia-smi every
    200 ms when the NVML binding is missing)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.halt, self.how = [], None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip()]
            phys = int(ids[index]) if ids and all(v.strip().isdigit() for v in ids) and index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.how = pynvml, "nvml every 10 ms"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi every 200 ms"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), 0), (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), 1),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), 2), (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), 3)]
        reasons_of = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.halt:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(reasons_of(self.handle))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                row = [str(sm), str(self.max_sm)] + ["Not Active"] * 4 + [str(pw)]
                for b, k in bits:
                    if mask & b:
                        row[2 + k] = "Active"
                self.rows.append((time.perf_counter(), row))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc and not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"]}
        if self.nvml:
            self.halt = True
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.25)
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = sorted({n for r in rows for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        pw = [float(r[6]) for r in rows if len(r) > 6 and r[6].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "power_w_max": max(pw) if pw else None, "how": self.how}


def run_reference_binary(n_phonons, outdir):
    """One run of the UNMODIFIED reference program on this workload; returns wall seconds."""
    from radiative3d_b200 import workloads
    os.makedirs(outdir, exist_ok=True)
    t = time.perf_counter()
    p = subprocess.run([REF_MAIN] + workloads.cmdline(WORKLOAD, n_phonons, TOA_DEGREE, outdir), cwd=outdir,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if p.returncode != 0:
        raise RuntimeError(f"reference binary failed with rc {p.returncode}")
    return time.perf_counter() - t


def cpu_baseline_one_core(sample=1_500_000):
    """The reference binary on ONE host core: (sample phonons) / (wall - model-build time)."""
    if not os.path.exists(REF_MAIN):
        return cpu_baseline_port(sample // 10, 1)
    with tempfile.TemporaryDirectory() as tmp:
        t_init = run_reference_binary(10, os.path.join(tmp, "init"))
        t_run = run_reference_binary(sample, os.path.join(tmp, "run"))
    return {"value": sample / max(t_run - t_init, 1e-9), "unit": "phonons/s", "cores": 1, "kind": "reference",
            "sample": f"{sample} phonons of {WORKLOAD} at TOA degree {TOA_DEGREE} through oracle/_ref/r3d_ref_main "
                      f"(unmodified reference, -O3), {t_run:.1f} s wall minus {t_init:.1f} s model build (N=10 run)"}


def cpu_baseline_port(sample, threads):
    """Fallback when the reference binary did not travel: the C oracle (a port), on `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from radiative3d_b200 import reference_host
    m = reference_host.build_model(WORKLOAD, TOA_DEGREE)
    ob.run(m, 0, 1000, SEED, nthreads=threads)
    t = time.perf_counter()
    ob.run(m, 0, sample, SEED, nthreads=threads)
    dt = time.perf_counter() - t
    return {"value": sample / dt, "unit": "phonons/s", "cores": threads, "kind": "port",
            "sample": f"{sample} phonons of {WORKLOAD} at TOA degree {TOA_DEGREE} through oracle/liboracle.so"}


# ---------------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """bench.py --impl reference: the reference's CPU implementation on all host cores (independent processes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = max(1, min(os.cpu_count() or 1, 32))
    per_proc = 400_000
    base = {"metric": "phonons traced/sec", "unit": "phonons/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "gpu_launches": 0,
            "config": {"workload": f"{WORKLOAD} (do-halfspace-nearsrc50.sh), TOA degree {TOA_DEGREE}",
                       "phonons_per_step": cores * per_proc, "note": "bounded sample of the GPU arm's batch"}}
    if not os.path.exists(REF_MAIN):
        cb = cpu_baseline_port(cores * 20000, cores)
        base.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": "phonons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(base))
        return 0
    from radiative3d_b200 import workloads
    with tempfile.TemporaryDirectory() as tmp:
        t_init = run_reference_binary(10, os.path.join(tmp, "init"))

        def step(i):       # (the reference seeds with time(NULL): equal seeds duplicate phonons, not cost)
            t = time.perf_counter()
            ps = []
            for c in range(cores):
                d = os.path.join(tmp, f"s{i}_{c}")
                os.makedirs(d, exist_ok=True)
                ps.append(subprocess.Popen([REF_MAIN] + workloads.cmdline(WORKLOAD, per_proc, TOA_DEGREE, d), cwd=d,
                                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            for p in ps:
                if p.wait() != 0:
                    raise RuntimeError("reference binary failed")
            return time.perf_counter() - t

        for i in range(min(args.warmup, 1)):          # one warm-up pass is enough for a CPU process farm
            step(-1 - i)
        times = [step(i) for i in range(args.steps)]
    sim = [max(t - t_init, 1e-9) for t in times]
    value = cores * per_proc * len(sim) / sum(sim)
    cb = {"value": value, "unit": "phonons/s", "cores": cores, "kind": "reference",
          "sample": f"{cores} independent processes x {per_proc} phonons per step through oracle/_ref/r3d_ref_main; "
                    f"wall of the slowest minus {t_init:.1f} s model build"}
    base.update(value=value, ms_per_step=1e3 * sum(sim) / len(sim), cpu_baseline=cb,
                e2e={"value": value, "unit": "phonons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(base))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--per-gpu", type=int, default=PER_GPU, help="phonons per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries ONE JSON line: whatever libraries print to fd 1 meanwhile (NCCL's version banner) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from radiative3d_b200 import abi, distributed, engine, reference_host
    from radiative3d_b200.model import _ARRAYS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the propagate path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE {world}; using {world}", file=sys.stderr)

    # ---- the model: reference host code builds it, we pin it ------------------------------------------------
    model = reference_host.build_model(WORKLOAD, TOA_DEGREE)
    keep = []
    for name, _ in _ARRAYS:
        t = torch.from_numpy(getattr(model, name)).pin_memory()
        keep.append(t)
        setattr(model, name, t.numpy())
    per, K, W = args.per_gpu, args.steps, args.warmup
    eng = engine.Engine(model, devices=(local,))

    def enqueue(i):
        first, n = distributed.shard_range(i * per * world, per * world, rank, world)
        eng.run_simulation(n, seed=SEED, first_phonon=first)

    de, dc, dk = (torch.as_tensor(v, device=f"cuda:{local}") for v in eng.device_accumulators(0))
    for i in range(W):
        enqueue(i)
    eng.sync()
    distributed.all_reduce_results(de, dc, dk)        # warm-up of the collective too (NCCL sets its channels up lazily)
    torch.cuda.synchronize()
    eng.reset()

    # ---- timed region: K steps + the one all-reduce of the bins ----------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        enqueue(W + i)
    dev_s = eng.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    distributed.all_reduce_results(de, dc, dk)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    dev_s += ev0.elapsed_time(ev1) * 1e-3
    launches = eng.launch_count - launches0 + (3 if world > 1 else 0)
    clocks = sampler.stop(t0, t1) if sampler else None
    tt = torch.tensor([dev_s, t1 - t0, float(launches)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dev_s, wall_s, launches = float(mx[0]), float(mx[1]), int(tt[2])
    else:
        wall_s = t1 - t0
    total = K * per * world
    counters = dk.cpu().numpy().astype(np.uint64)
    if int(counters[abi.R3D_CNT_PHONONS]) != total:
        raise SystemExit(f"bench.py: traced {int(counters[abi.R3D_CNT_PHONONS])} phonons, expected {total}")
    events_total = int(counters[abi.R3D_CNT_EVENTS])

    # ---- roofline of the propagate kernel (the only kernel of the path): CUDA events on its launching stream, live,
    # over one extra step of the same size (rank 0's GPU; the kernels are per GPU)
    roofline = None
    eng.reset()
    eng.set_profiling(True)
    enqueue(W + K)
    eng.sync()
    kt = eng.kernel_times()
    kt["phonons"] = per
    eng.set_profiling(False)
    if rank == 0:
        bytes_draw = algorithmic_bytes_per_draw(model.n_toa)
        alg = kt["draws"] * bytes_draw + kt["catches"] * BYTES_PER_CATCH
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj["dram_bytes_per_phonon"] * kt["phonons"] / max(kt["launches"], 1) if "phonons" in kt else None
            traffic_note = tj.get("source")
        except (OSError, ValueError, KeyError):
            pass
        sec, nl = kt["seconds"], max(kt["launches"], 1)
        achieved = alg / sec / 1e9
        ph = kt["phase1_seconds"] + kt["phase2_seconds"]
        roofline = {"bound": "hbm", "kernel": "propagate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "avg_launch_ms": 1e3 * sec / nl, "algorithmic_bytes_per_launch": alg / nl,
                    "algorithmic_bytes": f"{bytes_draw} B per table draw + {BYTES_PER_CATCH} B per bin update (SURVEY 8d)",
                    "share_of_step": {"propagate_kernel": 1.0},
                    "phase_share": {"advance+refill": kt["phase1_seconds"] / ph if ph else None,
                                    "draw+face": kt["phase2_seconds"] / ph if ph else None},
                    "units": {"phonons": per, "loop_events": kt["events"], "table_draws": kt["draws"], "bin_updates": kt["catches"]},
                    "ctas": kt["ctas"], "iterations_of_busiest_cta": kt["iterations"]}

    # ---- e2e: host model in, host bins out, through the C ABI ---------------------------------------------------
    eng.close()
    del de, dc, dk
    e2e_times = []
    h2d = model.table_bytes()
    d2h = model.n_seis * model.n_bins * (abi.R3D_BIN_NF64 * 8 + abi.R3D_BIN_NCNT * 8) + abi.R3D_NCOUNTERS * 8
    host_e = torch.zeros((model.n_seis, model.n_bins, abi.R3D_BIN_NF64), dtype=torch.float64).pin_memory()      # the caller's result buffers
    host_c = torch.zeros((model.n_seis, model.n_bins, abi.R3D_BIN_NCNT), dtype=torch.int64).pin_memory()
    host_out = (host_e.numpy(), host_c.numpy().view(np.uint64))
    for i in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        with engine.Engine(model, devices=(local,)) as e2:
            first, n = distributed.shard_range((W + K + 1 + i) * per * world, per * world, rank, world)
            e2.run_simulation(n, seed=SEED, first_phonon=first)
            e2.sync()
            ee, cc, kk = e2.fetch(out=host_out)
        dt = time.perf_counter() - t
        if i > 0:                                   # first pass warms the allocator / context
            e2e_times.append(dt)
    e2e_s = sum(e2e_times) / len(e2e_times) if e2e_times else float("nan")
    if world > 1:
        tmx = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        e2e_s = float(tmx[0])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_one_core()

    if rank == 0:
        out = {
            "metric": "phonons traced/sec", "value": total / dev_s, "unit": "phonons/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD} (do-halfspace-nearsrc50.sh): Halfspace model, 144 seismometers x 1250 bins, "
                                   f"TOA degree {TOA_DEGREE} ({model.n_toa} take-off angles, {h2d / 1e9:.2f} GB of tables)",
                       "phonons_per_gpu_per_step": per, "global_phonons_per_step": per * world, "parallelism": f"phonon-index sharding x{world}, "
                       "one NCCL all-reduce of the bins at the end",
                       "cache": "inputs larger than L2: 0.42 GB of CDF tables + 0.23 GB of guide tables, gathered at random, vs 126 MB L2",
                       "timing": "CUDA events on the launching stream (r3d_sync) + torch events around the all-reduce, max over ranks",
                       "wall_ms_per_step": 1e3 * wall_s / K,
                       "loop_events_per_second": events_total / dev_s, "events_per_phonon": events_total / total,
                       "model_built_by": "integration/_build/r3d_gpu_main (reference host code)"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": per * world / e2e_s, "unit": "phonons/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "r3d_create(pinned host model) + r3d_run + r3d_fetch(pinned host bins) + r3d_destroy, wall clock"},
            "roofline": roofline,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
