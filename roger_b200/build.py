"""In-tree build of libroger_b200.so (nvcc, sm_100a only).  Called by __graft_entry__.build().

Every .cu is compiled to its own object (in parallel, rebuilt only when it or a header it includes changed) and the objects
are linked into the shared library.  ROGER_NVCC_EXTRA adds flags (experiments: -DROGER_...=...), ROGER_B200_LIB redirects the
output; objects of different flag sets do not mix (the object directory carries a hash of the flags)."""
import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.environ.get("ROGER_B200_LIB") or os.path.join(LIBDIR, "libroger_b200.so")
INC = os.path.join("..", "..", "include")
COMMON_HEADERS = ["roger_pow.h", "roger_pow_tables.h", os.path.join(INC, "roger_b200.h"), os.path.join(INC, "roger_b200_fields.h"),
                  os.path.join(INC, "roger_b200_sas.h")]
SOURCES = {  # source -> headers of its own
    "svat_api.cu": ["svat_device.cuh"],
    "sas_api.cu": ["sas_device.cuh", "roger_gamma.h"],
}
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # NumPy rounds every operation: no FMA contraction of a*b+c in the physics (libdevice pow/exp keep their own fma)
    "-fmad=false",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _flags():
    extra = os.environ.get("ROGER_NVCC_EXTRA", "").split()
    return [f for f in NVCC_FLAGS if not (f == "-fmad=false" and "-fmad=true" in extra)] + extra


def _objdir():
    tag = hashlib.sha1(" ".join(_flags()).encode()).hexdigest()[:10]
    return os.path.join(HERE, "build", tag)


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _deps(src):
    return [os.path.join(CSRC, f) for f in [src] + SOURCES[src] + COMMON_HEADERS]


def needs_build():
    objs = [os.path.join(_objdir(), s.replace(".cu", ".o")) for s in SOURCES]
    return _stale(LIB, objs) or any(_stale(o, _deps(s)) for o, s in zip(objs, SOURCES))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    od = _objdir()
    os.makedirs(od, exist_ok=True)
    flags = _flags() + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(od, src.replace(".cu", ".o"))
        if force or _stale(obj, _deps(src)):
            subprocess.check_call([_nvcc()] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)])
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB
