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
