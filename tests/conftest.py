import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_goldens.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def shim():
    """The kernels' __host__ __device__ per-thread code compiled for the CPU (tests/host_shim; built once, cached)."""
    from tests import cpu_backend
    return cpu_backend.build_shim()


@pytest.fixture()
def cpu_backend(monkeypatch, shim):
    """install(loss) -> stand-in of the C ABI on top of the host shim (tests/cpu_backend.py), for one test."""
    from tests import cpu_backend as cb
    return cb.make_cpu_backend(monkeypatch, shim)
