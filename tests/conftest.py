import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 through gpurun)")


@pytest.fixture(scope="session")
def rtc():
    """The product package (ctypes host over librtc_b200.so); builds the library in-tree if it is stale or missing."""
    build = importlib.import_module("ray-tracer-challenge-rust_b200.build")
    build.build()
    return importlib.import_module("ray-tracer-challenge-rust_b200")


@pytest.fixture(scope="session")
def oracle(rtc):
    import helpers
    return helpers.load_oracle()


@pytest.fixture(scope="session")
def hostsim(rtc):
    import helpers
    return helpers.load_hostsim()


@pytest.fixture(scope="session")
def has_gpu(rtc):
    return rtc.device_count() > 0
