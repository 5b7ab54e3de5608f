"""ctypes mirror of include/roger_b200.h (structs, field ids).  No compute here."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_FIELDS_H = os.path.join(ROOT, "include", "roger_b200_fields.h")


def _parse_fields():
    txt = open(_FIELDS_H).read()
    body = txt[txt.index("#define ROGER_SVAT_FIELDS(X)"): txt.index("#define ROGER_FIELD_ENUM")]
    return re.findall(r"X\(([PSF]), (\w+)\)", body)


FIELDS = _parse_fields()  # [(kind, name)] in plane order
FIELD_NAMES = [n for _, n in FIELDS]
FIELD_KINDS = [k for k, _ in FIELDS]
FIELD_ID = {n: i for i, n in enumerate(FIELD_NAMES)}
N_FIELDS = len(FIELDS)
SLOTS_PER_DAY = 144


class SvatSettings(C.Structure):
    _fields_ = [
        ("ta_fm", C.c_double), ("hpi", C.c_double), ("end_event", C.c_double), ("sf", C.c_double),
        ("rmax", C.c_double), ("r_mp", C.c_double), ("l_sc", C.c_double), ("pi", C.c_double),
        ("transp_water_stress", C.c_double), ("dx", C.c_double), ("atol", C.c_double), ("rtol", C.c_double),
        ("enable_lateral_flow", C.c_int32), ("enable_macropore_lower_boundary_condition", C.c_int32),
        ("clamp_small_negative_storage", C.c_int32), ("forcing_uniform", C.c_int32),
        ("enable_groundwater_boundary", C.c_int32), ("reserved_", C.c_int32), ("kf_max", C.c_double),
    ]


class SvatScalars(C.Structure):
    _fields_ = [
        ("time", C.c_int64), ("itt", C.c_int64), ("dt_secs", C.c_int64), ("itt_day", C.c_int64),
        ("event_id", C.c_int64 * 2), ("event_id_counter", C.c_int64), ("time_event0", C.c_int64),
        ("month", C.c_int64), ("sanity_ok", C.c_int64), ("dt", C.c_double),
    ]


class SvatLayout(C.Structure):
    _fields_ = [
        ("ps_base", C.c_uint64), ("flux_base", C.c_uint64), ("n_tiles", C.c_int64),
        ("tile_cells", C.c_int32), ("n_ps_planes", C.c_int32), ("n_flux_planes", C.c_int32), ("reserved", C.c_int32),
    ]


def default_settings(lateral=False, forcing_uniform=True, **over):
    from .params import SETTINGS_DEFAULTS as D

    s = SvatSettings(
        ta_fm=D["ta_fm"], hpi=D["hpi"], end_event=D["end_event"], sf=D["sf"], rmax=D["rmax"], r_mp=D["r_mp"],
        l_sc=D["l_sc"], pi=D["pi"], transp_water_stress=D["transp_water_stress"], dx=D["dx"], atol=D["atol"],
        rtol=D["rtol"], enable_lateral_flow=int(lateral), enable_macropore_lower_boundary_condition=0,
        clamp_small_negative_storage=int(not lateral), forcing_uniform=int(forcing_uniform),
        enable_groundwater_boundary=0, reserved_=0, kf_max=D.get("kf_max", 3600.0),
    )
    for k, v in over.items():
        setattr(s, k, v)
    return s


def initial_scalars():
    """roger/variables.py initial values: dt=1, dt_secs=3600, event_id_counter=1, everything else 0."""
    s = SvatScalars()
    s.dt = 1.0
    s.dt_secs = 3600
    s.event_id_counter = 1
    s.sanity_ok = 1
    return s
