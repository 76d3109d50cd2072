import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import util  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    return util.Oracle()


@pytest.fixture(scope="session")
def sim():
    return util.Sim()


@pytest.fixture(scope="session")
def lib():
    """libvafgpu.so, built if needed (nvcc cross-compiles without a GPU)."""
    import subprocess
    subprocess.run(["make", "-s", "-C", util.PKG], check=True)
    return util.vafgpu.load_library()
