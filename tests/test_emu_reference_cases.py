"""A thinned-out run of the reference's known-answer tests on the host-emulated kernels."""

import pytest

from tests import reference_cases as rc
from tests.emu_support import emu_core_class

P = rc.TEST_PARAMS


@pytest.fixture(scope="module")
def dft():
    return emu_core_class()(P["W"], P["N"], P["xM_size"], P["yN_size"])


def test_emu_facet_to_subgrid_basic(dft):
    rc.facet_to_subgrid_basic(dft, 227, 415, thin=5)


def test_emu_facet_to_subgrid_dft_1d(dft):
    rc.facet_to_subgrid_dft_1d(dft, 228, 415, thin=23)


def test_emu_subgrid_to_facet_basic(dft):
    rc.subgrid_to_facet_basic(dft, 227, 416, thin=7)


def test_emu_subgrid_to_facet_dft(dft):
    rc.subgrid_to_facet_dft(dft, 228, 416, thin=9)


def test_emu_swift_configs_construct():
    """reference tests/test_core.py:82-90: every catalogue entry with N < 4096 constructs."""
    from ska_sdp_distributed_fourier_transform_b200 import SWIFT_CONFIGS

    n = 0
    for config in SWIFT_CONFIGS.values():
        if config["N"] < 4 * 1024:
            core = emu_core_class()(config["W"], config["N"], config["xM_size"], config["yN_size"])
            assert core.xM_yN_size == config["xM_size"] * config["yN_size"] // config["N"]
            n += 1
    assert n > 20
    assert len(SWIFT_CONFIGS) >= 244
