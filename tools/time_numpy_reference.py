#!/usr/bin/env python
"""Times the UNMODIFIED reference (NumPy backend, /root/reference) on its own benchmark workloads, in the build container,
with the reference harness' idiom (benchmarks/run_benchmarks.py:268-286: wall time per `RogerSetup.step`, the first three
iterations dropped).  The reference is pure Python and does not travel to the GPU box, so the record is committed
(profiles/numpy_reference_timing.json) and bench.py carries it in `cpu_baseline.numpy_reference`, saying where it was taken.

    python tools/time_numpy_reference.py            # ~2 minutes
"""
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import numpy as np  # noqa: E402

import ref_harness as rh  # noqa: E402
from roger_b200 import synth  # noqa: E402

BURN_IN = 3


def time_steps(model, nsteps):
    state = model.state
    times = []
    for _ in range(nsteps):
        t0 = time.perf_counter()
        model.step(state)
        times.append(time.perf_counter() - t0)
    return times[BURN_IN:]


def svat_case(side, lateral, hetero, ndays, doy):
    shape = (side, side)
    forc = synth.mixed_forcing(ndays, seed=42, start_doy=doy)
    if hetero:
        params, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lateral=lateral)
    else:
        params, th_rz, th_ss = synth.benchmark_parameters(shape, lateral=lateral)
    return dict(nx=side, ny=side, ndays=ndays, params=params, theta_rz0=th_rz, theta_ss0=th_ss, lateral=lateral, **forc)


def run_svat(name, side, lateral, hetero, nsteps):
    case = svat_case(side, lateral, hetero, 40, 121)
    model = rh.make_svat_setup(case)()
    model.setup()
    t = time_steps(model, nsteps + BURN_IN)
    cells = side * side
    return {"workload": name, "cells": cells, "steps_timed": len(t), "seconds_per_step": float(np.mean(t)),
            "cell_timesteps_per_s": cells / float(np.mean(t)), "min_seconds_per_step": float(np.min(t))}


def run_sas(side, ages, ndays):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_sas_golden as msg

    msg.CASES["_timing"] = dict(shape=(side, side), ndays=ndays + BURN_IN, ages=ages, doy=100,
                                sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5)))
    case = msg.build("_timing")
    model = rh.make_o18_setup(case)()
    model.setup()
    model.warmup()
    t = time_steps(model, ndays + BURN_IN)
    cells = side * side
    return {"workload": f"svat_oxygen18 offline SAS transport, power-law SAS, {ages} age bins, 6 sub-steps", "cells": cells, "ages": ages,
            "steps_timed": len(t), "seconds_per_step": float(np.mean(t)), "cell_timesteps_per_s": cells / float(np.mean(t))}


def main():
    rh.import_reference()
    out = {
        "what": "reference NumPy backend (unmodified /root/reference), float64, single process: wall time per RogerSetup.step, first 3 steps dropped",
        "where": "build container (the GPU box has no copy of the reference)", "host": platform.processor() or platform.machine(),
        "cores_available": os.cpu_count(), "cores_used": 1, "numpy": np.__version__, "python": platform.python_version(),
        "runs": [],
    }
    out["runs"].append(run_svat("SVAT_benchmark.py parameters (homogeneous), 100x100 cells, synthetic forcing from 1 May", 100, False, False, 40))
    out["runs"].append(run_svat("oneD model, heterogeneous parameters, 128x128 cells, synthetic forcing from 1 May", 128, True, True, 30))
    out["runs"].append(run_svat("oneD model, heterogeneous parameters, 512x512 cells, synthetic forcing from 1 May", 512, True, True, 8))
    out["runs"].append(run_sas(8, 1500, 4))
    out["runs"].append(run_sas(16, 1000, 3))
    path = os.path.join(ROOT, "profiles", "numpy_reference_timing.json")
    json.dump(out, open(path, "w"), indent=1)
    for r in out["runs"]:
        print("%-90s %10.4g cell-timesteps/s (%.3f s per step)" % (r["workload"], r["cell_timesteps_per_s"], r["seconds_per_step"]))


if __name__ == "__main__":
    main()
