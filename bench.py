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
        self.lines = []

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
            self.lines.append(line.strip())

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
        for ln in self.lines:
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

    cores = os.cpu_count() or 1
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
    print(json.dumps(line))


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
    for _ in range(args.warmup):
        m.step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    eng.kernel_ms = 0.0
    eng.launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    kinds = {}
    for _ in range(args.steps):
        sc = m.step(False)
        kinds[int(sc.dt_secs)] = kinds.get(int(sc.dt_secs), 0) + 1
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
    out_rz = np.empty(C)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        sc0 = eng.get_scalars()
        midnight = sc0.time % 86400 == 0
        sc = m.step(False)
        d2h += 4 + 88
        if midnight:
            h2d += 3 * 144 * 8
        if sc.time % 86400 == 0:
            out_rz[:] = m.get("theta_rz")
            _ = m.get("S")
            d2h += 2 * 8 * C
    eng.synchronize()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * C * args.steps / float(t[0])

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
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_cell_timestep": algo,
                         "kernel_ms_per_step": 1e3 * kern_time / args.steps},
            "e2e": {"value": e2e_value, "unit": "cell-timesteps/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    from oracle.engine import OracleSvatEngine

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--workload", choices=["svat", "oned"], default="oned")
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--ref-side", type=int, default=384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
