"""Executes the drop-in boundary: integration/roger_cuda_backend.patch is applied to a COPY of the reference package, the
copy is imported with ROGER_BACKEND=cuda, and the reference's SHIPPED SVAT model (roger/models/svat/svat.py, file IO
stubbed) is run through the reference's own `RogerSetup.setup()` / `RogerSetup.step()` -- which now routes the per-cell
work to roger_b200.integration.cuda_step -- against the same script on the unmodified NumPy backend.

Build container only (needs /root/reference; the container has no GPU, so the engine behind the cuda backend is the CPU
oracle here -- tests may plug it in; the GPU tests drive the same driver on the CUDA engine through a stand-in setup)."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "roger")), reason="reference checkout not present on this box")


def _run(pkg_parent, backend, out, mirror="eager", nx=4, ny=3, ndays=12):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([pkg_parent, ROOT]), ROGER_BACKEND=backend, ROGER_B200_TEST_ENGINE="oracle",
               ROGER_CUDA_MIRROR=mirror)
    env.pop("NPY_DISABLE_CPU_FEATURES", None)
    r = subprocess.run([sys.executable, os.path.join(HERE, "cuda_backend_runner.py"), out, str(nx), str(ny), str(ndays)], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    with np.load(out) as d:
        return {k: d[k] for k in d.files}


@pytest.fixture(scope="module")
def patched(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import apply_reference_patch

    target = str(tmp_path_factory.mktemp("roger_cuda"))
    apply_reference_patch.apply(REF, target)  # fails if the committed patch no longer applies
    return target


def test_patch_registers_the_backend(patched):
    src = open(os.path.join(patched, "roger", "backend.py")).read()
    assert 'BACKENDS = ("numpy", "jax", "cuda")' in src
    assert "cuda_step" in open(os.path.join(patched, "roger", "roger.py")).read()
    # the unpatched reference rejects the name
    r = subprocess.run([sys.executable, "-c", "import roger; print(roger.runtime_settings.backend)"], env=dict(os.environ, PYTHONPATH=REF, ROGER_BACKEND="cuda"), capture_output=True, text=True)
    assert r.returncode != 0  # "cuda" is not a member of its BACKENDS


def test_shipped_svat_model_on_the_cuda_backend_equals_the_numpy_backend(patched, tmp_path):
    ref = _run(patched, "numpy", str(tmp_path / "numpy.npz"))
    dut = _run(patched, "cuda", str(tmp_path / "cuda.npz"))
    assert str(dut["backend"]) == "cuda" and str(dut["driver"]) == "CudaStepDriver" and str(dut["engine"]) == "OracleSvatEngine"
    assert str(ref["driver"]) == ""
    assert ref["scalars"].shape == dut["scalars"].shape and ref["scalars"].shape[0] > 60
    assert np.array_equal(ref["scalars"], dut["scalars"])  # time, dt, itt_day, event ids ... of every step
    from tests.util import normalised_max_error

    for k in ref:
        if k in ("scalars", "backend", "driver", "engine", "downloads", "refused"):
            continue
        assert ref[k].shape == dut[k].shape, k
        # every recorded step, ghost cells and both time levels included
        assert normalised_max_error(ref[k], dut[k]) <= 1e-10, k


def test_lazy_mirrors_download_only_what_is_read(patched, tmp_path):
    eager = _run(patched, "cuda", str(tmp_path / "eager.npz"), mirror="eager", ndays=6)
    lazy = _run(patched, "cuda", str(tmp_path / "lazy.npz"), mirror="lazy", ndays=6)
    assert np.array_equal(eager["scalars"], lazy["scalars"])
    nsteps = eager["scalars"].shape[0]
    # the runner reads 23 variables after every step, 4 of them fluxes the lazy mirror refuses ... unless asked: it records
    # state variables only in lazy mode -- see below -- so the download count is (#recorded state variables) x steps
    assert int(lazy["downloads"]) <= 3 * nsteps + 8  # three variables read per step (+ the month change reading vs.ta ...)
    assert int(lazy["downloads"]) < int(eager["downloads"]) / 10
    assert int(lazy["refused"]) == 1  # an unrequested flux variable is refused loudly instead of being served stale
    from tests.util import normalised_max_error

    tau = 1
    for k in eager:
        if k in ("scalars", "backend", "driver", "engine", "downloads", "refused") or k not in lazy:
            continue
        a, b = eager[k], lazy[k]
        if a.ndim == 4:  # (step, x, y, level): the lazy mirror shows the step's result in both levels
            a, b = a[..., tau], b[..., tau]
        assert normalised_max_error(a[:, 2:-2, 2:-2], b[:, 2:-2, 2:-2]) <= 1e-10, k
    assert nsteps > 30
