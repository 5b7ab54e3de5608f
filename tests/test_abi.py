"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("roger_b200.h", "roger_b200_sas.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(roger_[a-z0-9_]+)\s*\(", txt))
    return sorted(n for n in names if not n.endswith("_fn"))


@pytest.fixture(scope="module")
def cuda_lib():
    from roger_b200 import build

    path = build.build()
    return C.CDLL(path)


def test_every_declared_entry_point_is_exported(cuda_lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(cuda_lib, s)]
    assert not missing, missing


def test_field_table_matches_header(cuda_lib):
    from roger_b200 import abi

    cuda_lib.roger_svat_field_name.restype = C.c_char_p
    n = cuda_lib.roger_svat_field_count()
    assert n == abi.N_FIELDS
    assert [cuda_lib.roger_svat_field_name(i).decode() for i in range(n)] == abi.FIELD_NAMES


def test_no_device_fails_loudly(cuda_lib):
    """Without a CUDA device the product path must raise, not fall back (this container has no GPU)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from roger_b200 import abi, sas
    from roger_b200.engine import CudaSvatEngine, RogerCudaError
    from roger_b200.sas_engine import CudaSasEngine

    with pytest.raises(RogerCudaError, match="no CUDA device"):
        CudaSvatEngine(16, abi.default_settings())
    with pytest.raises(RogerCudaError, match="no CUDA device"):
        CudaSasEngine(16, sas.default_settings(10))


def test_window_sums_of_the_hourly_aggregates_equal_the_full_row_sums():
    """the kernels read six forcing slots per hourly aggregate instead of walking the 144-slot row with 138 zeros: same bits
    (including NaN slots, negative zeros, windows across the two halves of numpy's pairwise reduction and at the day's end)"""
    import ctypes as C

    import numpy as np

    from roger_b200.build import LIB

    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    lib.roger_selftest_window_sums.argtypes = [dp, C.c_int, dp]
    rng = np.random.default_rng(5)
    out = np.zeros(4)
    for trial in range(40):
        row = rng.normal(0, 3, 144) * (rng.random(144) < 0.7)
        row[rng.random(144) < 0.1] = -0.0
        if trial % 3 == 0:
            row[rng.random(144) < 0.15] = np.nan
        for itt_day in range(144):
            assert lib.roger_selftest_window_sums(row.ctypes.data_as(dp), itt_day, out.ctypes.data_as(dp)) == 0
            a, b, c, d = out.view(np.uint64)
            nan_sum = np.isnan(out[0]) and np.isnan(out[1])
            assert a == b or nan_sum, (trial, itt_day, out)
            assert c == d or (np.isnan(out[2]) and np.isnan(out[3])), (trial, itt_day, out)
