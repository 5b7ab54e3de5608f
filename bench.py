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
# DRAM bytes per cell and launch of the dominant kernel from the committed `ncu --set full` captures (profiles/r1_ncu_summary.md:
# dram__bytes_read.sum + dram__bytes_write.sum divided by the cells of the captured launch); None where no capture exists
NCU_DRAM_BYTES_PER_CELL = {"oned": 1122.0, "svat": None}  # r1f capture: 4.707 GB per launch at 2048^2
NCU_DRAM_BYTES_PER_BIN_DAY = 1327.0  # k_sas_day, per cell and age bin


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
                                   "power-law SAS (SVATOXYGEN18_benchmark.py), 6 sub-steps"},
            "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
                             "sample": f"{cells} cells x {args.ages} age bins of the same workload, {args.steps} daily steps, OpenMP over cells"},
            "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return
    side = args.ref_side
    shape = (side, side)
    m = build_model(OracleSvatEngine, shape, args.workload, seed=43)
    for _ in range(args.warmup):
        m.step(False)
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
        "config": {"workload": f"{args.workload} {args.side}x{args.side} cells per GPU (oneD_benchmark.py), uniform synthetic forcing from 1 May"},
        "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
                         "sample": f"{side}x{side} cells of the same workload, {args.steps} adaptive steps, OpenMP over cells"},
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
    reducer = None
    if world > 1:
        from roger_b200.parallel import make_flag_reducer

        reducer = make_flag_reducer(dist, torch.device("cuda", local))
    m = build_model(CudaSvatEngine, shape, args.workload, seed=43 + rank, engine_kwargs={"device": local, "flag_reducer": reducer})
    eng = m.engine

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
    ms_by_dt.clear()
    n_by_dt.clear()
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
    # Steady-state placement of the window: a daily slab leaves the device asynchronously while the NEXT day's steps run, so a
    # window that happens to end a few steps after a midnight would time the tail of a copy a long run never waits for.  The
    # (untimed) steps below move the window's start to noon; the window still holds one midnight per simulated day.
    for _ in range(48):
        if eng.get_scalars().time % 86400 == 12 * 3600:
            break
        m.run_steps(1)
    e2e_start = "noon"
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    left = args.steps
    while left > 0:
        sc0 = eng.get_scalars()
        if sc0.time % 86400 == 0:
            h2d += 3 * 144 * 8
        done = m.run_steps(left)
        left -= done
        sc = eng.get_scalars()
        d2h += (4 + 88) * done
        if sc.time % 86400 == 0:
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

    sas_res = None
    if not args.no_sas:
        # second half of the BASELINE metric: SVAT + offline oxygen-18 SAS transport (configs[2]) on the same GPU(s)
        eng.close()
        del m
        torch.cuda.empty_cache()
        sas_res = run_sas_cuda(args, local, world, rank, dist if world > 1 else None, min(args.steps, args.sas_steps), 3, args.sas_side, args.ages)

    if rank == 0:
        peak, peak_src = measured_peak()
        algo = ALGO_BYTES[args.workload]
        ach = algo * C * args.steps / kern_time / 1e9
        line = {
            "metric": "cell-timesteps/s", "value": value, "unit": "cell-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_time / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload} {side}x{side} cells per GPU (oneD_benchmark.py), uniform synthetic forcing from 1 May",
                "cells_per_gpu": C, "step_mix_dt_secs": kinds, "l2": "state planes are %.1f GB per GPU, far larger than the 126 MB L2" % (C * 8 * 120 / 1e9),
                "parallelism": f"cell-block x{world}",
            },
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": (NCU_DRAM_BYTES_PER_CELL[args.workload] * C if NCU_DRAM_BYTES_PER_CELL[args.workload] else None),
                         "traffic_source": "profiles/r1_ncu_summary.md (bytes per cell of the captured launch x cells per launch)",
                         "peak_source": peak_src, "algorithmic_bytes_per_cell_timestep": algo,
                         "kernel_ms_per_step": 1e3 * kern_time / args.steps,
                         "kernel_ms_per_step_by_dt_secs": {k: v / n_by_dt[k] for k, v in ms_by_dt.items() if n_by_dt.get(k)}},
            "e2e": {"value": e2e_value, "unit": "cell-timesteps/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "window_start": e2e_start},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
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
                     "traffic": NCU_DRAM_BYTES_PER_BIN_DAY * ages * C, "traffic_frac_of_peak": NCU_DRAM_BYTES_PER_BIN_DAY * ages * C * steps / kern / 1e9 / peak,
                     "traffic_source": "profiles/r1_ncu_summary.md (bytes per cell and age bin of the captured launch x cells x bins)",
                     "peak_source": peak_src, "algorithmic_bytes_per_cell_timestep": algo, "kernel_ms_per_step": 1e3 * kern / steps,
                     "note": "exact-order solver streams its two scratch arrays once per sub-step: measured DRAM traffic is ~20x the "
                             "algorithmic bytes (profiles/), see DESIGN.md"},
        "e2e": {"value": world * C * steps / e2e_wall, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 9 * 8 * C,
                "d2h_bytes_per_step": len(sas.OUTPUTS) * 8 * C},
        "gpu_launches": int(launches), "finite_C_rz_fraction": finite,
    }


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
            "sample": f"{cells} cells x {ages} age bins of the same workload, {days} daily steps, OpenMP over cells"}


def cpu_baseline(args):
    from oracle.engine import OracleSvatEngine

    use_all_host_cores()
    side = args.ref_side
    m = build_model(OracleSvatEngine, (side, side), args.workload, seed=43)
    for _ in range(2):
        m.step(False)
    n = max(10, min(args.steps, 40))
    t0 = time.perf_counter()
    for _ in range(n):
        m.step(False)
    el = time.perf_counter() - t0
    return {"value": side * side * n / el, "unit": "cell-timesteps/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{side}x{side} cells of the same workload, {n} adaptive steps, OpenMP over cells"}


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
