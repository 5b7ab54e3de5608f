"""Restart files of the cuda backend: the run state a later process needs to continue bit-identically -- every parameter and
state plane of the SVAT / oneD engine with the scheduler's scalars and the host-side cursors of the hooks, or the age-ranked
arrays of a transport engine -- written where the reference writes its restart file (roger/restart.py:32-174: all variables with
`write_to_restart`, plus the scalars).  Container: NumPy's `.npz` (the reference's is HDF5 through h5py, which this image does
not have; the layout is one array per variable under the reference's variable name, so a converter is a loop).

    model.run(...); restart.write_svat(model, "SVAT.restart.npz")
    model2 = SvatModel(...same inputs...); model2.setup(); restart.read_svat(model2, "SVAT.restart.npz"); model2.run(...)
"""
import numpy as np

from . import abi

_SCALARS = ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "month", "sanity_ok", "dt")


def write_svat(model, path):
    """roger.restart.write_restart for the online SVAT / oneD models (roger/restart.py:130-174)"""
    eng = model.engine
    kinds = dict(zip(abi.FIELD_NAMES, abi.FIELD_KINDS))
    out = {"plane_" + n: eng.download(n) for n in abi.FIELD_NAMES if kinds[n] in "PS"}
    sc = eng.get_scalars()
    for k in _SCALARS:
        out["scalar_" + k] = np.asarray(getattr(sc, k))
    out["scalar_event_id"] = np.asarray([sc.event_id[0], sc.event_id[1]], dtype=np.int64)
    out["host_itt_forc"] = np.asarray(model.itt_forc)
    out["host_month"] = np.asarray(model.month, dtype=np.int64)
    out["host_n_steps"] = np.asarray(model.n_steps)
    out["shape"] = np.asarray(model.shape, dtype=np.int64)
    out["lateral"] = np.asarray(int(model.lateral))
    np.savez_compressed(path, **out)


def read_svat(model, path):
    """roger.restart.read_restart (roger/restart.py:32-67): overwrites the planes and scalars of a model that was set up with the
    same grid; the next forcing day is handed over again at the next midnight (a restart file is written at one)."""
    with np.load(path) as d:
        if tuple(int(v) for v in d["shape"]) != tuple(model.shape) or bool(int(d["lateral"])) != bool(model.lateral):
            raise ValueError("restart file belongs to a different grid / model")
        eng = model.engine
        for k in d.files:
            if k.startswith("plane_"):
                eng.upload(k[len("plane_"):], d[k])
        sc = eng.get_scalars()
        for k in _SCALARS:
            setattr(sc, k, d["scalar_" + k].item())
        sc.event_id[0], sc.event_id[1] = int(d["scalar_event_id"][0]), int(d["scalar_event_id"][1])
        eng.set_scalars(sc)
        model.itt_forc = int(d["host_itt_forc"])
        model.month = [int(v) for v in d["host_month"]]
        model.n_steps = int(d["host_n_steps"])
    if int(sc.time) % 86400 != 0:
        raise ValueError("restart files are written at midnight (the forcing day block is not part of the file)")


def write_transport(model, path, names=None):
    """the age-ranked arrays of a transport model (SasModel / NitrateModel) and its host-side cursors"""
    names = list(names or getattr(model, "RESTART_STATE", None) or ["sa_rz", "sa_ss", "msa_rz", "msa_ss"])
    out = {"state_" + n: model.state(n) for n in names}
    out["host_itt"] = np.asarray(model.itt)
    out["host_C_snow"] = np.asarray(model.C_snow)
    out["ages"] = np.asarray(model.ages)
    np.savez_compressed(path, **out)


def read_transport(model, path):
    with np.load(path) as d:
        if int(d["ages"]) != int(model.ages):
            raise ValueError("restart file belongs to a different age discretisation")
        order = ["sa_rz", "sa_ss", "msa_rz", "msa_ss", "Nmin_rz", "Nmin_ss"]
        for k in d.files:
            if k.startswith("state_"):
                name = k[len("state_"):]
                if name in ("Nmin_rz", "Nmin_ss"):
                    model.engine.upload_state_index(order.index(name), d[k])
                else:
                    model.engine.upload_state(name, d[k])
        model.itt = int(d["host_itt"])
        model.C_snow = np.array(d["host_C_snow"])
