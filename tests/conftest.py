import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _emu_call(flat, batch, afd_capacity=0):
    from tests import emu
    return emu.call_batch(flat, batch, afd_capacity=afd_capacity)


def _cuda_call(flat, batch, afd_capacity=0):
    from varlociraptor_b200 import engine  # raises without libvlr_engine.so; vlr_ctx_create fails without a device
    eng = engine.PosteriorEngine(flat)
    out = eng.call_batch(batch, afd_capacity=afd_capacity)
    assert batch.n_loci == 0 or eng.launches >= 1
    return out


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def engine_call(request):
    """The engine under test as `call(flat_scenario, batch, afd_capacity=0) -> CallResults`: the single-lane host build
    of the kernel source (tests/emu, control flow only, runs without a GPU) and the CUDA library through the C-ABI
    (-m gpu). The same test body checks both against the oracle."""
    fn = _emu_call if request.param == "emu" else _cuda_call
    fn.kind = request.param
    return fn
