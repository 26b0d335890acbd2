import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle_backend import load_oracle
    return load_oracle()


# Worlds that are not the particles bundle run on the generic one-launch program, which exists twice: the interpreter
# kernel (csrc/generic_program.cuh, schema read from the parameter block, tile in shared memory) and the registration's own
# kernel compiled by NVRTC at bgr_build (csrc/generic_program_jit.cuh, schema as compile-time constants, rows in
# registers).  GPU test modules of the generic path opt in with `pytestmark = pytest.mark.usefixtures("generic_kernel")`
# and then run every test on both.
@pytest.fixture(params=["interpreter", "jit", "jit_quarter_tiles"])
def generic_kernel(request, monkeypatch):
    monkeypatch.setenv("BGR_TUNE_JIT", "0" if request.param == "interpreter" else "2")
    if request.param == "jit":
        monkeypatch.setenv("BGR_TUNE_JIT_ITEM", "512")   # one tile per work item (what worlds of many tiles per SM run)
    # "jit_quarter_tiles": the default for worlds of few tiles per SM — 128-row work items
    return request.param
