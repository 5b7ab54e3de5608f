"""In-tree build of libroger_b200.so (nvcc, sm_100a only).  Called by __graft_entry__.build()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.environ.get("ROGER_B200_LIB") or os.path.join(LIBDIR, "libroger_b200.so")
SOURCES = ["svat_api.cu", "sas_api.cu"]
HEADERS = ["svat_device.cuh", "sas_device.cuh", "roger_pow.h", "roger_pow_tables.h", os.path.join("..", "..", "include", "roger_b200.h"),
           os.path.join("..", "..", "include", "roger_b200_fields.h"), os.path.join("..", "..", "include", "roger_b200_sas.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # NumPy rounds every operation: no FMA contraction of a*b+c in the physics (libdevice pow/exp keep their own fma)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("ROGER_NVCC_EXTRA", "").split()
    flags = [f for f in NVCC_FLAGS if not (f == "-fmad=false" and "-fmad=true" in extra)]
    cmd = [_nvcc()] + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.check_call(cmd)
    return LIB
