"""The regularised lower incomplete gamma function of the gamma SAS family (roger_b200/csrc/roger_gamma.h; the reference calls
scipy.special.gammainc, roger/core/sas.py:153) against a committed table of scipy outputs (tests/golden/make_gammainc_table.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.engine import lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ABS_TOL, REL_TOL = 2e-14, 1e-12  # relative: prefactor x^a e^-x / Gamma(a) in the far tails (p ~ 1e-160 at a = 170)


def _table():
    with np.load(os.path.join(GOLDEN, "gammainc_table.npz")) as d:
        return {k: d[k] for k in d.files}


def _oracle(a, x, mode):
    L = lib()
    L.sas_oracle_gammainc.restype = C.c_double
    L.sas_oracle_gammainc.argtypes = [C.c_double, C.c_double]
    L.sas_oracle_set_pow(mode)
    try:
        return np.array([L.sas_oracle_gammainc(float(ai), float(xi)) for ai, xi in zip(a, x)])
    finally:
        L.sas_oracle_set_pow(0)


def _check(mine, t):
    err = np.abs(mine - t["p"])
    assert err.max() <= ABS_TOL, err.max()
    ok = t["p"] > 1e-300
    assert (err[ok] / t["p"][ok]).max() <= REL_TOL
    small = t["a"] <= 30  # the shape parameters a SAS function plausibly takes
    assert (err[ok & small] / t["p"][ok & small]).max() <= 1e-13


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_gammainc_against_scipy_table(mode):
    """mode 0: C library pow / exp (the mode the oracle is pinned in); mode 1: roger_pow / roger_exp (what the device evaluates)"""
    t = _table()
    _check(_oracle(t["a"], t["x"], mode), t)


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_gammainc_argument_rules(mode):
    t = _table()
    mine = _oracle(t["edge_a"], t["edge_x"], mode)
    assert np.array_equal(np.isnan(mine), np.isnan(t["edge_p"]))
    ok = ~np.isnan(mine)
    assert np.array_equal(mine[ok], t["edge_p"][ok])


@pytest.mark.gpu
def test_device_gammainc_equals_host_bit_for_bit_and_scipy_table():
    from roger_b200 import build

    L = C.CDLL(build.build())
    dp = C.POINTER(C.c_double)
    L.roger_gammainc_device.argtypes = [dp, dp, dp, C.c_int64, C.c_int]
    t = _table()
    a = np.ascontiguousarray(np.concatenate([t["a"], t["edge_a"]]))
    x = np.ascontiguousarray(np.concatenate([t["x"], t["edge_x"]]))
    out = np.empty_like(a)
    assert L.roger_gammainc_device(a.ctypes.data_as(dp), x.ctypes.data_as(dp), out.ctypes.data_as(dp), a.size, 0) == 0
    host = _oracle(a, x, 1)
    assert np.array_equal(out.view(np.uint64)[~np.isnan(host)], host.view(np.uint64)[~np.isnan(host)])
    assert np.array_equal(np.isnan(out), np.isnan(host))
    _check(out[:t["a"].size], t)
