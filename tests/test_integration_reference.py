"""The drop-in boundary: an UNMODIFIED RogerSetup subclass of the reference is
stepped by roger_b200.integration.CudaStepDriver and must reproduce what the reference's own `RogerSetup.step` computes
-- same hooks, same `vs` arrays with their tau / taum1 levels, same scalars.

Runs only where the reference checkout is importable (the build container); the engine plugged in here is the CPU
oracle (tests may do that), the product default is the CUDA engine, which has no fallback."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/roger"), reason="reference checkout not present on this box")


def _case(name):
    import make_svat_golden as mg

    case, _ = mg.build_case(name)
    return case


@pytest.mark.parametrize("name,nsteps", [("svat_mixed_dist", 260), ("oned_rain", 50), ("svat_snowrain", 50)])
def test_driver_reproduces_reference_step(name, nsteps):
    import ref_harness as rh
    from oracle.engine import OracleSvatEngine
    from roger_b200 import abi
    from roger_b200.integration import CudaStepDriver
    from tests.util import normalised_max_error

    case = _case(name)
    Setup = rh.make_svat_setup(case)
    ref = Setup()
    ref.setup()
    dut = Setup()
    dut.setup()
    driver = CudaStepDriver(dut, engine_cls=OracleSvatEngine)
    names = [n for n in abi.FIELD_NAMES if n in driver._active and n != "maskRiverLake"]
    worst = {}
    for step in range(nsteps):
        ref.step(ref.state)
        driver.step()
        a, b = ref.state.variables, dut.state.variables
        for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt"):
            assert float(np.asarray(getattr(a, k))) == float(np.asarray(getattr(b, k))), (step, k)
        assert list(a.event_id) == list(b.event_id), step
        if step % 7 == 0 or step == nsteps - 1:
            for n in names:
                x, y = np.asarray(getattr(a, n)), np.asarray(getattr(b, n))
                assert x.shape == y.shape and x.dtype == y.dtype
                if n.endswith("_num_error"):
                    continue
                e = normalised_max_error(x[2:-2, 2:-2], y[2:-2, 2:-2])  # both time levels where the variable has them
                if e > worst.get(n, (0, 0))[0]:
                    worst[n] = (e, step)
    bad = {n: v for n, v in worst.items() if v[0] > 1e-10}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0])[:10]


def test_driver_refuses_offline_transport_and_has_no_fallback():
    import ref_harness as rh
    from roger_b200.engine import RogerCudaError
    from roger_b200.integration import CudaStepDriver

    Setup = rh.make_svat_setup(_case("svat_rain"))
    m = Setup()
    m.setup()
    import torch

    if not torch.cuda.is_available():
        d = CudaStepDriver(m)  # product default engine
        with pytest.raises(RogerCudaError, match="no CUDA device"):
            d.step()
