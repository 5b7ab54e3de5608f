"""Restart files of the cuda backend (roger_b200/restart.py; the reference: roger/restart.py:32-174): a run continued from a
restart file is bit-identical to the uninterrupted run -- the criterion of the reference's own test/restart_test.py.  CPU: the
oracle engines behind the same host-side models; GPU: the CUDA engines."""
import numpy as np
import pytest

from roger_b200 import abi, restart, sas, synth
from roger_b200.svat import SvatModel
from tests import util


def _svat_roundtrip(engine_cls, tmp_path, lateral):
    shape = (9, 7)
    forcing = synth.mixed_forcing(30, seed=42, start_doy=100)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lateral=lateral, masked_fraction=0.1)
    weights = synth.distributed_weights(shape, seed=44)

    def make():
        m = SvatModel(engine_cls, shape, forcing, prim, th_rz, th_ss, lateral=lateral, weights=weights)
        m.setup()
        return m

    a = make()
    a.run(6 * 86400)
    path = str(tmp_path / "SVAT.restart.npz")
    restart.write_svat(a, path)
    a.run(5 * 86400)
    b = make()
    restart.read_svat(b, path)
    b.run(5 * 86400)
    sa, sb = a.engine.get_scalars(), b.engine.get_scalars()
    assert (sa.time, sa.itt, sa.event_id_counter, sa.time_event0) == (sb.time, sb.itt, sb.event_id_counter, sb.time_event0)
    kinds = dict(zip(abi.FIELD_NAMES, abi.FIELD_KINDS))
    for n in abi.FIELD_NAMES:
        if kinds[n] in "PS":
            assert util.same_bits(a.get(n), b.get(n)), n


def _sas_roundtrip(engine_cls, tmp_path):
    from tests import util_sas as u

    fx = u.load_case("sas_power_snow")
    a = u.model_from_fixture(fx, engine_cls)
    for _ in range(9):
        a.step()
    path = str(tmp_path / "O18.restart.npz")
    restart.write_transport(a, path)
    b = u.model_from_fixture(fx, engine_cls)
    restart.read_transport(b, path)
    for _ in range(8):
        a.step()
        b.step()
    for k in sas.STATE:
        assert u.same(a.state(k), b.state(k)), k
    oa, ob = a.outputs(), b.outputs()
    for k in oa:
        assert u.same(oa[k], ob[k]), k


@pytest.mark.parametrize("lateral", [False, True])
def test_svat_run_continued_from_a_restart_file_is_bit_identical(tmp_path, lateral):
    from oracle.engine import OracleSvatEngine

    _svat_roundtrip(OracleSvatEngine, tmp_path, lateral)


def test_transport_run_continued_from_a_restart_file_is_bit_identical(tmp_path):
    from oracle.engine import OracleSasEngine

    _sas_roundtrip(OracleSasEngine, tmp_path)


@pytest.mark.gpu
def test_restart_files_on_the_cuda_engines(tmp_path):
    from roger_b200.engine import CudaSvatEngine
    from roger_b200.sas_engine import CudaSasEngine

    _svat_roundtrip(CudaSvatEngine, tmp_path, True)
    _sas_roundtrip(CudaSasEngine, tmp_path)
