import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build librbp_b200.so and the oracle when a compiler is present (the GPU box uses the prebuilt files)."""
    import shutil

    from robopoker_b200 import build

    have_lib = os.path.exists(build.LIB) and os.path.exists(build.ORACLE_LIB)
    if shutil.which("g++") and os.path.exists(build.NVCC):
        try:
            build.build_lib()
            build.build_oracle()
        except Exception:
            if not have_lib:
                raise
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding

    return binding


@pytest.fixture(scope="session")
def rbp():
    import robopoker_b200

    return robopoker_b200
