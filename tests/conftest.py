"""pytest configuration: markers, import paths and shared fixtures."""

import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_1d():
    return numpy.load(os.path.join(GOLDEN, "ref_1d_n1024.npz"))


@pytest.fixture(scope="session")
def golden_2d():
    return numpy.load(os.path.join(GOLDEN, "ref_2d_n256.npz"))


@pytest.fixture(scope="session")
def golden_windows():
    return numpy.load(os.path.join(GOLDEN, "ref_windows.npz"))
