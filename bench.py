#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the fused SVAT / oneD step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA backend (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the CPU oracle port on the host cores

A "step" is one adaptive time step (RogerSetup.step, roger/roger.py:355-489) over every cell of the
grid.  N=1 workload: BASELINE.json configs[1] -- the oneD model (Green-Ampt infiltration, percolation,
lateral flow) on 4096x4096 cells, spatially uniform synthetic forcing (the reference benchmark
broadcasts one series, benchmarks/SVAT_benchmark.py:162-170), heterogeneous seeded parameters.
For N>1 every rank holds its own 4096x4096 block (weak scaling); the only exchange is the 1-word flag
all-reduce of the adaptive time stepping.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from roger_b200 import abi, synth  # noqa: E402
from roger_b200.svat import SvatModel  # noqa: E402

# algorithmic bytes per cell-timestep (2*state + parameters + forcing)
ALGO_BYTES = {"svat": 984.0, "oned": 1200.0}


def ncu_traffic():
    """DRAM bytes of the dominant kernels as `ncu --set full` measured them (dram__bytes_read.sum + dram__bytes_write.sum of the
    captured launch divided by its cells / cells x bins): read from profiles/ncu_traffic.json, which tools/ncu_traffic.py writes
    from the committed captures; absent entries give traffic = null"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}


def numpy_reference_record(workload, ages=None):
    """the unmodified reference's NumPy backend timed in the build container (tools/time_numpy_reference.py; the reference is
    pure Python and does not travel to the GPU box)"""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "numpy_reference_timing.json")))
    except Exception:
        return None
    key = "SAS" if workload == "sas" else ("oneD" if workload == "oned" else "SVAT_benchmark")
    runs = [r for r in rec["runs"] if key in r["workload"]]
    if ages is not None and any(r.get("ages") == ages for r in runs):
        runs = [r for r in runs if r.get("ages") == ages]
    if not runs:
        return None
    best = max(runs, key=lambda r: r["cell_timesteps_per_s"])
    return {"value": best["cell_timesteps_per_s"], "unit": "cell-timesteps/s", "cores": rec["cores_used"], "kind": "reference (NumPy backend)",
            "sample": best["workload"] + f", {best['cells']} cells, {best['steps_timed']} steps", "where": rec["where"], "host": rec["host"],
            "source": "profiles/numpy_reference_timing.json"}


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs time the oracle on every host core and say how many"""
    from oracle import engine

    n = os.cpu_count() or 1
    engine.lib().roger_oracle_set_threads(n)
    return n


def sas_algo_bytes(ages):
    """each of sa_rz, sa_ss, msa_rz, msa_ss read and written once per day + forcing + scalar outputs"""
    return 64.0 * ages + 344.0
START_DOY = 121  # 1 May: rain events, hourly and daily steps, no snow


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []  # (arrival time, text)
        self.t_mark = 0.0

    def mark(self):
        """start of the timed region: only samples from here on are reported (earlier ones are the fallback for regions
        shorter than the sampling period)"""
        self.t_mark = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = [ln for t, ln in self.lines if t >= self.t_mark]
        if not timed:
            timed = [ln for _, ln in self.lines[-3:]]
        for ln in timed:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons)}


def build_model(engine_cls, shape, workload, seed, engine_kwargs=None, ndays=40):
    lateral = workload == "oned"
    forcing = synth.mixed_forcing(ndays, seed=42, start_doy=START_DOY)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=seed, lateral=lateral)
    m = SvatModel(engine_cls, shape, forcing, prim, th_rz, th_ss, lateral=lateral, engine_kwargs=engine_kwargs)
    m.setup()
    return m


def place_window(m):
    """Untimed steps up to the midnight that opens the first day stepped hourly -- the day the first rain event of the synthetic
    series arrives -- so that the timed window holds rain steps as well as dry ones (both arms use it).  Whether a day is
    stepped hourly is only known after its first step, hence the engine's in-memory checkpoint / rollback.  Returns the day."""
    eng = m.engine
    for _ in range(40):
        while eng.get_scalars().time % 86400 != 0:
            m.run_steps(144)
        m.checkpoint()
        m.run_steps(1)
        hourly = eng.get_scalars().dt_secs == 3600
        m.rollback()
        if hourly:
            break
        m.run_steps(1)
    return int(eng.get_scalars().time // 86400)


def run_reference(args):
    """The reference arm: the reference's CPU algorithm (oracle port, all host threads) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.engine import OracleSvatEngine

    cores = use_all_host_cores()
    if args.workload == "sas":
        from oracle.engine import OracleSasEngine
        from roger_b200 import sas

        cells = args.ref_cells
        daily = sas_daily_tile(OracleSvatEngine, args.steps + args.warmup + 2)
        eng = OracleSasEngine(cells, sas.default_settings(args.ages))
        model = SasBenchModel(eng, cells, args.ages, daily)
        for _ in range(args.warmup):
            eng.step(model.day_forcing())
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.step(model.day_forcing())
        el = time.perf_counter() - t0
        v = cells * args.steps / el
        emit({
            "impl": "reference", "metric": "cell-timesteps/s", "value": v, "unit": "cell-timesteps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"svat_oxygen18 offline SAS transport {args.sas_side}x{args.sas_side} cells per GPU, {args.ages} age bins, "
                                   "power-law SAS (SVATOXYGEN18_benchmark.py), 6 sub-steps",
                       "sample": f"timed on a {cells}-cell sample of that workload, {cores} host threads"},
            "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
                             "sample": f"{cells} cells x {args.ages} age bins of the same workload, {args.steps} daily steps, OpenMP over cells",
                             "numpy_reference": numpy_reference_record("sas", args.ages)},
            "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return
    side = args.ref_side
    shape = (side, side)
    m = build_model(OracleSvatEngine, shape, args.workload, seed=43)
    for _ in range(args.warmup):
        m.step(False)
    day = place_window(m)  # the same simulated window as the CUDA arm
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.step(False)
    el = time.perf_counter() - t0
    C = side * side
    v = C * args.steps / el
    line = {
        "impl": "reference", "metric": "cell-timesteps/s", "value": v, "unit": "cell-timesteps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} {args.side}x{args.side} cells per GPU (oneD_benchmark.py), uniform synthetic forcing from 1 May",
                   "sample": f"timed on a {side}x{side}-cell sample of that workload (same parameter and forcing generators), {cores} host threads",
                   "window": "starts at the midnight opening simulated day %d (first rain event of the series)" % day},
        "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
                         "sample": f"{side}x{side} cells of the same workload, {args.steps} adaptive steps, OpenMP over cells",
                         "numpy_reference": numpy_reference_record(args.workload)},
        "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_cuda(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from roger_b200.engine import CudaSvatEngine
    from roger_b200.parallel import bind_to_gpu_numa_node

    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)  # page-locked buffers are first-touched on the GPU's own NUMA node

    if args.workload == "sas":
        sampler = ClockSampler(local)
        sampler.start()
        res = run_sas_cuda(args, local, world, rank, dist if world > 1 else None, args.steps, max(args.warmup, 1), args.sas_side, args.ages,
                           sampler=sampler)
        clocks = sampler.stop()
        if rank == 0:
            line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": world, "steps": res["steps"],
                    "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(res["config"], parallelism=f"cell-block x{world}"),
                    "roofline": res["roofline"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": clocks}
            if world == 1 and not args.no_cpu_baseline:
                os.sched_setaffinity(0, affinity0)  # the CPU leg runs on every host core
                line["cpu_baseline"] = sas_cpu_baseline(args.ages, cells=args.ref_cells)
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    side = args.side
    shape = (side, side)
    C = side * side
    m = build_model(CudaSvatEngine, shape, args.workload, seed=43 + rank, engine_kwargs={"device": local})
    eng = m.engine
    if world > 1:  # the one data-path exchange: ncclAllGather of the scheduler's flag words on the engine's stream
        eng.enable_nccl_flags(dist, torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, device-timed ---------------------------------------------
    ms_by_dt, n_by_dt = {}, {}

    def advance(n):
        """exactly n adaptive steps (day-blocked launches where the scheduler allows); returns {dt_secs: count}"""
        kinds, left = {}, n
        while left > 0:
            # the engine batches pre-schedulable steps (one launch per batch on small grids, queued single-step launches on
            # large ones); last_step_dts() tells which steps they were
            ms0 = eng.kernel_ms
            done = m.run_steps(left)
            dts = eng.last_step_dts()
            for d in dts:
                kinds[d] = kinds.get(d, 0) + 1
            if len(set(dts)) == 1:
                ms_by_dt[dts[0]] = ms_by_dt.get(dts[0], 0.0) + (eng.kernel_ms - ms0)
                n_by_dt[dts[0]] = n_by_dt.get(dts[0], 0) + done
            left -= done
        return kinds

    sampler = ClockSampler(local)
    sampler.start()
    advance(args.warmup)
    # Placement of the timed window (untimed steps, place_window()): the midnight that opens the first day stepped hourly.  In
    # `e2e` the daily output of the day that just ended is issued at that midnight, inside the timed region, and its copy to the
    # host overlaps the day's steps: the steady state of a long run (one output per simulated day).
    window_start_day = place_window(m)
    ms_by_dt.clear()
    n_by_dt.clear()
    m.checkpoint()  # `e2e` below repeats exactly this simulated window (same state, same forcing days, same step mix)
    barrier()
    sampler.mark()
    eng.kernel_ms = 0.0
    eng.launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    kinds = advance(args.steps)
    eng.synchronize()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    kernel_ms = eng.kernel_ms
    launches = eng.launches
    t = torch.tensor([wall, kernel_ms / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_time, kern_time = float(t[0]), float(t[1])
    value = world * C * args.steps / step_time

    # ---- e2e: public API, host buffers, H2D / D2H inside the timed region -------------------------
    # per step: scalars + flags read back (always); at each midnight: the day's forcing block from
    # host memory and a D2H of the daily output slab (theta_rz, S) as the reference's diagnostics would write.
    pin = [torch.empty(C, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
    eng.reserve_snapshots(2)  # what a driver does when it configures its output: nothing is allocated inside the stepping loop
    m.rollback()
    e2e_start = "the window `value` was timed on (device-resident checkpoint, rolled back): midnight opening simulated day %d" % window_start_day
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    left = args.steps
    # the output of the day that ended at this midnight (the slabs the reference's diagnostics write once per day)
    e2e_output = os.environ.get("ROGER_BENCH_E2E_OUTPUT", "1") != "0"  # (experiment switch: e2e without the daily slabs)
    if e2e_output:
        eng.download_async("theta_rz", pin[0])
        eng.download_async("S", pin[1])
        d2h += 2 * 8 * C
    while left > 0:
        sc0 = eng.get_scalars()
        if sc0.time % 86400 == 0:
            h2d += 3 * 144 * 8
        done = m.run_steps(left)
        left -= done
        sc = eng.get_scalars()
        d2h += (4 + 88) * done
        if sc.time % 86400 == 0 and e2e_output:
            eng.download_wait()  # the previous day's slabs have left the pinned buffers (the writer thread owns copies)
            eng.download_async("theta_rz", pin[0])
            eng.download_async("S", pin[1])
            d2h += 2 * 8 * C
    eng.download_wait()
    eng.synchronize()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * C * args.steps / float(t[0])

    eng.close()
    del m
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_sweep:
        extras["sweep"], extras["strong"] = run_sweep(args, local, world, rank, dist if world > 1 else None, value, step_time / args.steps)
    if not args.no_dropin and rank == 0:
        extras["dropin"] = run_dropin(args, local)
    if not args.no_scenarios:
        extras["scenarios"] = {"snow_month_per_cell_forcing": run_scenario_snow_percell(args, local, world, rank, dist if world > 1 else None)}

    if not args.no_nitrate:
        extras["nitrate"] = run_nitrate_cuda(args, local, world, rank, dist if world > 1 else None)

    sas_res = None
    if not args.no_sas:
        # second half of the BASELINE metric: SVAT + offline oxygen-18 SAS transport (configs[2]) on the same GPU(s)
        sas_res = run_sas_cuda(args, local, world, rank, dist if world > 1 else None, min(args.steps, args.sas_steps), 3, args.sas_side, args.ages)

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic = ncu_traffic().get(args.workload)
        algo = ALGO_BYTES[args.workload]
        ach = algo * C * args.steps / kern_time / 1e9
        line = {
            "metric": "cell-timesteps/s", "value": value, "unit": "cell-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_time / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload} {side}x{side} cells per GPU (oneD_benchmark.py), uniform synthetic forcing from 1 May",
                "cells_per_gpu": C, "step_mix_dt_secs": kinds, "window": "starts at the midnight opening simulated day %d (first rain event of the series)" % window_start_day,
                "l2": "state planes are %.1f GB per GPU, far larger than the 126 MB L2" % (C * 8 * 120 / 1e9),
                "parallelism": f"cell-block x{world}",
            },
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": (traffic["dram_bytes_per_cell"] * C if traffic else None),
                         "traffic_source": (traffic or {}).get("source"),
                         "peak_source": peak_src, "algorithmic_bytes_per_cell_timestep": algo,
                         "kernel_ms_per_step": 1e3 * kern_time / args.steps,
                         "kernel_ms_per_step_by_dt_secs": {k: v / n_by_dt[k] for k, v in ms_by_dt.items() if n_by_dt.get(k)}},
            "e2e": {"value": e2e_value, "unit": "cell-timesteps/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "window_start": e2e_start},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(extras)
        if sas_res is not None:
            line["sas"] = sas_res
        line["e2e"]["host_numa_binding"] = numa
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, affinity0)  # the CPU leg runs on every host core
            line["cpu_baseline"] = cpu_baseline(args)
            if sas_res is not None:
                line["sas"]["cpu_baseline"] = sas_cpu_baseline(args.ages, cells=args.ref_cells)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _max_over_ranks(values, dist):
    import torch

    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def tiled_model(engine_cls, n_cells, workload, seed, engine_kwargs, tile=65536, start_doy=START_DOY, weights=False):
    """a grid of n_cells whose parameters repeat a heterogeneous tile of `tile` cells (BASELINE configs[4]: grids of 2^16 ..
    2^28 cells; drawing 2^28 independent parameter sets on the host would take longer than the measurement)"""
    lateral = workload == "oned"
    t = min(tile, n_cells)
    prim, th_rz, th_ss = synth.heterogeneous_parameters((t, 1), seed=seed, lateral=lateral)
    reps = -(-n_cells // t)
    rep = lambda a: np.tile(np.asarray(a).reshape(-1), reps)[:n_cells]  # noqa: E731
    prim = {k: rep(v) for k, v in prim.items()}
    w = None
    if weights:
        w = {k: rep(v) for k, v in synth.distributed_weights((t, 1), seed=seed + 1).items()}
    forcing = synth.mixed_forcing(60, seed=42, start_doy=start_doy)
    m = SvatModel(engine_cls, (n_cells, 1), forcing, prim, rep(th_rz), rep(th_ss), lateral=lateral, weights=w, engine_kwargs=engine_kwargs)
    m.setup()
    return m


def run_sweep(args, local, world, rank, dist, main_value, main_s_per_step):
    """BASELINE configs[4]: the same step on grids of 2^16 .. 2^24 cells per GPU (weak: every rank holds that many) and, on
    several GPUs, 2^24 cells IN TOTAL cut into contiguous blocks (strong).  2^26 and 2^27 cells per GPU with --sweep-large."""
    import torch
    from roger_b200.engine import CudaSvatEngine

    main_cells = args.side * args.side
    sizes = [1 << 16, 1 << 18, 1 << 20, 1 << 22] + ([1 << 26, 1 << 27] if args.sweep_large else [])
    strong_cells = (1 << 24) // world
    out = []

    def measure(n_cells):
        m = tiled_model(CudaSvatEngine, n_cells, args.workload, seed=43 + rank, engine_kwargs={"device": local})
        if dist is not None:
            m.engine.enable_nccl_flags(dist, torch.device("cuda", local))
        steps = min(args.steps, 24)
        left = args.warmup
        while left > 0:
            left -= m.run_steps(left)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        m.engine.kernel_ms = 0.0
        t0 = time.perf_counter()
        left = steps
        while left > 0:
            left -= m.run_steps(left)
        m.engine.synchronize()
        wall, kern = _max_over_ranks([time.perf_counter() - t0, m.engine.kernel_ms / 1e3], dist)
        m.engine.close()
        del m
        torch.cuda.empty_cache()
        return {"cells_per_gpu": n_cells, "cells_total": n_cells * world, "value": world * n_cells * steps / wall, "ms_per_step": 1e3 * wall / steps,
                "kernel_ms_per_step": 1e3 * kern / steps, "steps": steps}

    for n in sizes:
        out.append(measure(n))
    out.append({"cells_per_gpu": main_cells, "cells_total": main_cells * world, "value": main_value, "ms_per_step": 1e3 * main_s_per_step,
                "steps": args.steps, "note": "the headline measurement"})
    out.sort(key=lambda r: r["cells_per_gpu"])
    r = measure(strong_cells)  # the same tiled workload at every N, so the points of a 1/2/4/8 series are comparable
    strong = dict(r, cells_total=strong_cells * world, note=f"2^24 cells in total, {strong_cells} per GPU (fixed total work; tiled parameters, window right after warm-up)")
    return out, strong


def run_dropin(args, local):
    """Throughput of the reference-facing path itself: `roger_b200.integration.CudaStepDriver` stepping a setup object with the
    reference's conventions ((nx+4, ny+4[, 2]) host arrays with ghost cells and tau / taum1 levels, the SVAT model's hooks
    set_forcing / set_parameters / after_timestep; tests/mini_setup.py stands in for the reference package, which does not
    travel to the GPU box).  `eager`: every state and flux plane written back into vs.* after every step and the setup's
    after_timestep hook run on them -- what a user sees without changing anything; `lazy`: dirty-flag mirrors, planes are
    downloaded when read (here: `S` and `theta_rz` once per simulated day, as a daily diagnostic would)."""
    import torch
    from roger_b200.engine import CudaSvatEngine
    from roger_b200.integration import CudaStepDriver
    from tests.mini_setup import MiniSvatSetup

    side = args.dropin_side
    shape = (side, side)
    n_cells = side * side
    forcing = synth.mixed_forcing(40, seed=42, start_doy=START_DOY)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43)
    out = {"cells": n_cells, "setup": "tests/mini_setup.py::MiniSvatSetup (the reference's array conventions and SVAT hooks), SVAT model, uniform forcing"}
    for mode, steps in (("eager", 4), ("lazy", 30)):
        setup = MiniSvatSetup(shape, forcing, prim, th_rz, th_ss)
        setup.setup()
        drv = CudaStepDriver(setup, engine_cls=CudaSvatEngine, engine_kwargs={"device": local}, mirror=mode)
        for _ in range(2):
            drv.step()
        torch.cuda.synchronize()
        d0 = drv.downloads
        per_step, midnight = [], []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            at_midnight = int(np.asarray(setup.state.variables.time)) % 86400 == 0  # the setup's set_forcing assembles the day blocks
            sc = drv.step()
            if mode == "lazy" and sc.time % 86400 == 0:
                _ = setup.state.variables.S, setup.state.variables.theta_rz
            drv.engine.synchronize()
            (midnight if at_midnight else per_step).append(time.perf_counter() - t1)
        el = time.perf_counter() - t0
        out[mode] = {"value": n_cells * steps / el, "unit": "cell-timesteps/s", "ms_per_step": 1e3 * el / steps, "steps": steps,
                     "plane_downloads_per_step": (drv.downloads - d0) / steps,
                     "ms_per_step_within_a_day": 1e3 * float(np.median(per_step)) if per_step else None,
                     "ms_per_step_at_midnight": 1e3 * float(np.median(midnight)) if midnight else None,
                     "note": "at midnight the setup's own set_forcing hook assembles three (nx+4, ny+4, 144) host arrays and the driver hands them over"}
        drv.engine.close()
        del drv, setup
        torch.cuda.empty_cache()
    return out


def run_scenario_snow_percell(args, local, world, rank, dist):
    """A winter month with snow on the ground and per-cell forcing (Eberbaechle pattern: series x per-cell weights): no step
    can be scheduled ahead -- every adaptive step needs the grid-wide flags of its own forcing aggregates (k_aggregate_a) and
    of the snow state -- so every step pays two flag exchanges across the ranks.  The day blocks are built on the device
    (inputs resident in HBM, as for `value`); only stepping is timed."""
    import torch
    from roger_b200.engine import CudaSvatEngine

    side = args.scenario_side
    n_cells = side * side
    dev = torch.device("cuda", local)
    m = tiled_model(CudaSvatEngine, n_cells, args.workload, seed=143 + rank, engine_kwargs={"device": local}, start_doy=15, weights=True)
    eng = m.engine
    if dist is not None:
        eng.enable_nccl_flags(dist, dev)
    w = {k: torch.from_numpy(v).to(dev) for k, v in m.weights.items()}
    series = {k: torch.from_numpy(np.asarray(m.forcing[k], dtype=np.float64)).to(dev) for k in ("PREC", "TA", "PET")}
    keep = []

    def set_forcing(sc):  # SvatModel._set_forcing with the day blocks assembled on the device
        if sc.time % 86400 != 0:
            return
        sc.itt_day = 0
        i0 = m.itt_forc
        m.month[1] = int(m.forcing["MONTH"][i0])
        sc.month = m.month[1]
        sl = slice(i0, i0 + 144)
        blocks = [(series["PREC"][sl][None, :] * w["prec_weight"][:, None]).contiguous(),
                  (series["TA"][sl][None, :] + w["ta_offset"][:, None]).contiguous(),
                  (series["PET"][sl][None, :] * w["pet_weight"][:, None]).contiguous()]
        torch.cuda.synchronize()
        keep[:] = blocks
        eng.set_forcing_day_device(*[b.data_ptr() for b in blocks])
        m.itt_forc += 144

    m._set_forcing = set_forcing
    steps, warm = min(args.steps, 24), max(args.warmup, 3)
    kinds = {}
    for _ in range(warm):
        m.run_steps(1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    eng.kernel_ms, eng.launches = 0.0, 0
    wall = 0.0
    for _ in range(steps):
        sc = eng.get_scalars()
        m._set_forcing(sc)  # untimed: the day block is an input
        eng.set_scalars(sc)
        m._set_parameters(sc)
        t0 = time.perf_counter()
        eng.step(False)
        wall += time.perf_counter() - t0
        m.month[0] = m.month[1]
        for d in eng.last_step_dts():
            kinds[d] = kinds.get(d, 0) + 1
    snow = float((eng.download("swe") > 0).mean())
    wall, kern = _max_over_ranks([wall, eng.kernel_ms / 1e3], dist)
    launches = eng.launches
    eng.close()
    del m
    torch.cuda.empty_cache()
    return {"value": world * n_cells * steps / wall, "unit": "cell-timesteps/s", "ms_per_step": 1e3 * wall / steps, "kernel_ms_per_step": 1e3 * kern / steps,
            "steps": steps, "cells_per_gpu": n_cells, "step_mix_dt_secs": kinds, "fraction_of_cells_with_snow_at_end": snow,
            "gpu_launches": int(launches), "flag_exchanges_per_step": 2 if world > 1 else 0,
            "config": f"{args.workload} {side}x{side} cells per GPU, forcing from 15 January x per-cell weights, day blocks assembled on the device"}


# ---------------------------------------------------------------------------------------------------------------
# SVAT + offline oxygen-18 StorAge-selection transport (BASELINE.json configs[2]: power-law SAS, 1024x1024 cells)
SAS_SPEC = dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))  # SVATOXYGEN18_benchmark.py:129-138
SAS_TILE = (64, 64)


def sas_daily_tile(svat_engine_cls, ndays, engine_kwargs=None):
    """Daily hydrologic forcing of the transport model: an online SVAT run (the product's own SVAT path on the GPU arm,
    the oracle on the CPU arm) of a heterogeneous 64x64 tile, aggregated per day as the reference's
    test/make_data_for_svat_transport.py:279-300 does.  The tile is repeated over the transport grid."""
    from roger_b200 import sas_inputs

    forcing = synth.mixed_forcing(ndays + 2, seed=42, start_doy=START_DOY)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(SAS_TILE, seed=44)
    m = SvatModel(svat_engine_cls, SAS_TILE, forcing, prim, th_rz, th_ss, engine_kwargs=engine_kwargs)
    m.setup()
    daily = sas_inputs.daily_fluxes_from_svat(m, ndays)
    if hasattr(m.engine, "close"):
        m.engine.close()
    rng = np.random.default_rng(45)
    daily["C_ISO_IN"] = rng.uniform(-12, -5, daily["PREC"].shape)
    return daily


class SasBenchModel:
    """The driver loop of the offline transport benchmark over a tiled grid: per day, the host-side `set_forcing`
    (roger_b200.sas.SasModel._set_forcing on the tile) expanded to all cells, one engine step."""

    def __init__(self, engine, n_cells, ages, daily_tile):
        from roger_b200 import sas

        self.sas = sas
        self.engine = engine
        self.C = n_cells
        nt = daily_tile["PREC"].shape[0]
        self.tile = sas.SasModel(_NullEngine, nt, ages, daily_tile, SAS_SPEC)
        self.idx = np.arange(n_cells) % nt
        engine.set_parameters(np.ones(n_cells), sas.sas_parameter_block(n_cells, SAS_SPEC))
        if hasattr(engine, "init_linear"):
            engine.init_linear(daily_tile["S_RZ"][self.idx, 0], daily_tile["S_SS"][self.idx, 0], float(sas.delta_to_conc(-10.0)))
        else:
            c0 = float(sas.delta_to_conc(-10.0))
            msa = np.full((n_cells, ages), c0)
            msa[:, 0] = 0
            engine.upload_state("sa_rz", sas.initial_age_distribution(daily_tile["S_RZ"][self.idx, 0], ages))
            engine.upload_state("sa_ss", sas.initial_age_distribution(daily_tile["S_SS"][self.idx, 0], ages))
            engine.upload_state("msa_rz", msa)
            engine.upload_state("msa_ss", msa.copy())

    def day_forcing(self):
        """host planes (9, C) of the next day"""
        self.tile.itt += 1
        f = self.tile._set_forcing()
        self.tile.C_snow[:, 0] = self.tile.C_snow[:, 1]
        return np.ascontiguousarray(f[:, self.idx])


class _NullEngine:
    def __init__(self, *a, **k):
        pass


def run_sas_cuda(args, local, world, rank, dist, steps, warmup, side, ages, sampler=None):
    """Returns the `sas` result dict (value / roofline / e2e) measured on this rank's GPU, max over ranks."""
    import torch
    from roger_b200 import sas
    from roger_b200.engine import CudaSvatEngine
    from roger_b200.sas_engine import CudaSasEngine

    C = side * side
    ndays = 2 * (steps + warmup) + 1
    daily = sas_daily_tile(CudaSvatEngine, ndays, engine_kwargs={"device": local})
    eng = CudaSasEngine(C, sas.default_settings(ages), device=local)
    model = SasBenchModel(eng, C, ages, daily)
    host_days = [model.day_forcing() for _ in range(ndays - 1)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # value: forcing planes already resident in HBM
    dev_days = [torch.from_numpy(f).to(f"cuda:{local}") for f in host_days[: steps + warmup]]
    for d in range(warmup):
        eng.step_device(dev_days[d].data_ptr())
    barrier()
    if sampler is not None:
        sampler.mark()
    eng.kernel_ms = 0.0
    eng.launches = 0
    t0 = time.perf_counter()
    for d in range(warmup, warmup + steps):
        eng.step_device(dev_days[d].data_ptr())
    eng.synchronize()
    wall = time.perf_counter() - t0
    kernel_ms, launches = eng.kernel_ms, eng.launches
    del dev_days
    # e2e: host forcing planes in, per-cell result planes (22 concentrations / delta values) out, every day
    pinned = [torch.from_numpy(f).pin_memory() for f in host_days[steps + warmup: 2 * steps + warmup]]
    out = torch.empty((len(sas.OUTPUTS), C), dtype=torch.float64).pin_memory().numpy()
    barrier()
    t0 = time.perf_counter()
    for f in pinned:
        eng.step(f.numpy())
        eng.download_outputs(out)
    eng.synchronize()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([wall, kernel_ms / 1e3, e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, kern, e2e_wall = (float(v) for v in t)
    peak, peak_src = measured_peak()
    sas_traffic = ncu_traffic().get("sas")
    algo = sas_algo_bytes(ages)
    ach = algo * C * steps / kern / 1e9
    finite = float(np.isfinite(out[sas.OUTPUTS.index("C_rz")]).mean())
    eng.close()
    return {
        "metric": "cell-timesteps/s", "value": world * C * steps / wall, "unit": "cell-timesteps/s", "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * wall / steps,
        "config": {"workload": f"svat_oxygen18 offline SAS transport {side}x{side} cells per GPU, {ages} age bins, power-law SAS "
                               "(SVATOXYGEN18_benchmark.py), 6 sub-steps, daily forcing from an online SVAT run of a 64x64 tile",
                   "cells_per_gpu": C, "ages": ages,
                   "l2": "age-ranked state is %.1f GB per GPU, far larger than the 126 MB L2" % (C * ages * 8 * 4 / 1e9)},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": (sas_traffic["dram_bytes_per_bin_day"] * ages * C if sas_traffic else None),
                     "traffic_frac_of_peak": (sas_traffic["dram_bytes_per_bin_day"] * ages * C * steps / kern / 1e9 / peak if sas_traffic else None),
                     "traffic_source": (sas_traffic or {}).get("source"),
                     "peak_source": peak_src, "algorithmic_bytes_per_cell_timestep": algo, "kernel_ms_per_step": 1e3 * kern / steps,
                     "note": "exact-order solver streams its two scratch arrays once per sub-step: measured DRAM traffic is ~20x the "
                             "algorithmic bytes (profiles/), see DESIGN.md"},
        "e2e": {"value": world * C * steps / e2e_wall, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 9 * 8 * C,
                "d2h_bytes_per_step": len(sas.OUTPUTS) * 8 * C},
        "gpu_launches": int(launches), "finite_C_rz_fraction": finite,
    }


def run_nitrate_cuda(args, local, world, rank, dist):
    """BASELINE configs[3]: nitrate StorAge-selection transport (anion transport + soil nitrogen cycle on age-ranked mineral
    nitrogen) on an Eberbaechle-like grid at 5 m resolution, 400 x 265 = 106 000 cells per GPU, 1500 age bins; daily hydrologic
    forcing from an online SVAT run of the CUDA path on a 64 x 64 tile.  No collective: every rank steps its own block."""
    import torch
    from roger_b200 import nitrate, sas
    from roger_b200.engine import CudaSvatEngine
    from roger_b200.sas_engine import CudaSasEngine

    C, A = args.nitrate_cells, args.ages
    steps, warm = min(args.steps, 6), 2
    ndays = steps + warm + 1
    tile = sas_daily_tile(CudaSvatEngine, ndays, engine_kwargs={"device": local})
    nt = tile["PREC"].shape[0]
    idx = np.arange(C) % nt
    rng = np.random.default_rng(77 + rank)
    t = ndays + 1
    daily = {k: np.ascontiguousarray(v[idx]) for k, v in tile.items() if getattr(v, "ndim", 0) == 2 and v.shape[0] == nt}
    forcing = synth.mixed_forcing(ndays + 2, seed=42, start_doy=START_DOY)
    daily["TA"], daily["DOY"] = forcing["TA"][::144][:t].copy(), forcing["DOY"][::144][:t].astype(np.int64)
    daily["C_IN"] = rng.uniform(0, 60, (C, t)) * (rng.random((C, t)) < 0.1)
    daily["Z_ROOT"] = np.repeat(rng.uniform(200, 900, (C, 1)), t, axis=1)
    daily["LU_ID"] = np.repeat(rng.choice([8.0, 525.0, 583.0, 541.0], (C, 1)), t, axis=1)
    daily["NUP"] = rng.uniform(0, 80, (C, t))
    params = {k: rng.uniform(lo, hi, C) for k, (lo, hi) in dict(
        kfix_rz=(10, 60), kmin_rz=(20, 120), kmin_ss=(5, 40), km_nit_rz=(5, 40), dmax_nit_rz=(20, 200), kngl_rz=(1, 20), km_denit_rz=(5, 40),
        dmax_denit_rz=(10, 150), km_denit_ss=(5, 40), dmax_denit_ss=(5, 60), z_soil=(1200, 2500), damp_soil_temp=(12, 16)).items()}
    params["phi_soil_temp"] = np.full(C, 91.0)
    params["S_sat_rz"] = daily["S_RZ"][:, 0] * rng.uniform(1.02, 1.7, C)
    params["S_sat_ss"] = daily["S_SS"][:, 0] * rng.uniform(1.02, 1.7, C)
    anion = dict(alpha_transp=rng.uniform(0.2, 0.9, C), alpha_q=rng.uniform(0.4, 1.0, C), lu_id=daily["LU_ID"][:, 0], S_sat_rz=params["S_sat_rz"])
    spec = dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.25), transp=(6, 0.5), q_rz=(6, 2.0), q_ss=(6, 3.0))  # svat_crop_nitrate tutorial :255-268
    m = nitrate.NitrateModel(CudaSasEngine, C, A, daily, spec, params, anion, nitt=t, engine_kwargs={"device": local})
    m.setup()
    sa_rz, sa_ss = sas.initial_age_distribution(daily["S_RZ"][:, 0], A), sas.initial_age_distribution(daily["S_SS"][:, 0], A)
    decay = np.exp(-np.linspace(0.001, 7, A))[None, :]
    for k, arr in enumerate([sa_rz, sa_ss, 30.0 * sa_rz, 15.0 * sa_ss, 600.0 * decay * np.ones((C, 1)), 200.0 * decay * np.ones((C, 1))]):
        m.engine.upload_state_index(k, arr)
    del sa_rz, sa_ss
    for _ in range(warm):
        m.step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    m.engine.kernel_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        m.step()  # host forcing planes in (9 + 4), scalar results read back below: the public path
    out = m.nitrate_outputs()
    m.engine.synchronize()
    wall, kern = _max_over_ranks([time.perf_counter() - t0, m.engine.kernel_ms / 1e3], dist)
    finite = float(np.isfinite(out["Nmin_s"]).mean())
    m.engine.close()
    return {"value": world * C * steps / wall, "unit": "cell-timesteps/s", "ms_per_step": 1e3 * wall / steps, "kernel_ms_per_step": 1e3 * kern / steps,
            "kernel_value": world * C * steps / kern, "steps": steps, "cells_per_gpu": C, "ages": A, "finite_Nmin_s_fraction": finite,
            "config": f"nitrate SAS transport (anion transport + soil nitrogen cycle), {C} cells per GPU (Eberbaechle-like 400 x 265 grid at 5 m), "
                      f"{A} age bins, power-law SAS, 6 sub-steps; host forcing in, per-cell results out every day"}


def sas_cpu_baseline(ages, cells=8192, days=4):
    from oracle.engine import OracleSasEngine, OracleSvatEngine
    from roger_b200 import sas

    use_all_host_cores()
    daily = sas_daily_tile(OracleSvatEngine, days + 2)
    eng = OracleSasEngine(cells, sas.default_settings(ages))
    model = SasBenchModel(eng, cells, ages, daily)
    eng.step(model.day_forcing())
    t0 = time.perf_counter()
    for _ in range(days):
        eng.step(model.day_forcing())
    el = time.perf_counter() - t0
    return {"value": cells * days / el, "unit": "cell-timesteps/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{cells} cells x {ages} age bins of the same workload, {days} daily steps, OpenMP over cells",
            "numpy_reference": numpy_reference_record("sas", ages)}


def cpu_baseline(args):
    from oracle.engine import OracleSvatEngine

    use_all_host_cores()
    side = args.ref_side
    m = build_model(OracleSvatEngine, (side, side), args.workload, seed=43)
    for _ in range(2):
        m.step(False)
    place_window(m)  # the same simulated window as the CUDA measurement
    n = max(10, min(args.steps, 40))
    t0 = time.perf_counter()
    for _ in range(n):
        m.step(False)
    el = time.perf_counter() - t0
    return {"value": side * side * n / el, "unit": "cell-timesteps/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{side}x{side} cells of the same workload, {n} adaptive steps, OpenMP over cells",
            "numpy_reference": numpy_reference_record(args.workload)}


def _claim_stdout():
    """Libraries (NCCL's version banner, launch notices) write to file descriptor 1; the contract is ONE JSON line there.
    Everything else is sent to stderr and the line is written to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


_JSON_OUT = None


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--workload", choices=["svat", "oned", "sas"], default="oned")
    ap.add_argument("--sas-side", type=int, default=1024, help="SAS transport grid side per GPU (configs[2]: 1024)")
    ap.add_argument("--ages", type=int, default=1500, help="age bins of the SAS transport (BASELINE: ~1500)")
    ap.add_argument("--sas-steps", type=int, default=4, help="timed daily steps of the secondary SAS measurement")
    ap.add_argument("--ref-cells", type=int, default=8192, help="cells of the CPU sample of the SAS workload")
    ap.add_argument("--no-sas", action="store_true", help="skip the secondary SVAT + oxygen-18 SAS measurement")
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--ref-side", type=int, default=768, help="grid side of the CPU sample of the SVAT / oneD workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the grid-size sweep (2^16 .. 2^22 cells per GPU) and the strong-scaling point")
    ap.add_argument("--sweep-large", action="store_true", help="add 2^26 and 2^27 cells per GPU to the sweep")
    ap.add_argument("--no-nitrate", action="store_true", help="skip the nitrate transport measurement (BASELINE configs[3])")
    ap.add_argument("--nitrate-cells", type=int, default=400 * 265)
    ap.add_argument("--no-dropin", action="store_true", help="skip the throughput measurement of the drop-in driver (CudaStepDriver)")
    ap.add_argument("--dropin-side", type=int, default=1024)
    ap.add_argument("--no-scenarios", action="store_true", help="skip the snow-month / per-cell-forcing scenario")
    ap.add_argument("--scenario-side", type=int, default=2048)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
