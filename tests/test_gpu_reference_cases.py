"""The reference's tests/test_core.py known-answer tests against the CUDA backend."""

import pytest

from tests import reference_cases as rc

pytestmark = pytest.mark.gpu
P = rc.TEST_PARAMS


@pytest.fixture(scope="module")
def dft():
    from ska_sdp_distributed_fourier_transform_b200 import SwiftlyCoreB200

    return SwiftlyCoreB200(P["W"], P["N"], P["xM_size"], P["yN_size"])


def test_base_params(dft):
    """tests/test_core.py:43-79."""
    from ska_sdp_distributed_fourier_transform_b200 import SwiftlyCoreB200

    assert (dft.W, dft.N, dft.yN_size, dft.xM_size) == (P["W"], P["N"], P["yN_size"], P["xM_size"])
    assert dft.xM_yN_size == 128
    with pytest.raises(ValueError):
        SwiftlyCoreB200(P["W"], 1050, P["xM_size"], P["yN_size"])


SIZES = [(228, 416), (227, 416), (228, 415), (227, 415)]


@pytest.mark.parametrize("xA_size,yB_size", SIZES)
def test_facet_to_subgrid_basic(dft, xA_size, yB_size):
    rc.facet_to_subgrid_basic(dft, xA_size, yB_size)


@pytest.mark.parametrize("xA_size,yB_size", SIZES)
def test_facet_to_subgrid_dft_1d(dft, xA_size, yB_size):
    rc.facet_to_subgrid_dft_1d(dft, xA_size, yB_size)


def test_facet_to_subgrid_dft_2d(dft):
    rc.facet_to_subgrid_dft_2d(dft)


@pytest.mark.parametrize("xA_size,yB_size", SIZES)
def test_subgrid_to_facet_basic(dft, xA_size, yB_size):
    rc.subgrid_to_facet_basic(dft, xA_size, yB_size)


@pytest.mark.parametrize("xA_size,yB_size", SIZES)
def test_subgrid_to_facet_dft(dft, xA_size, yB_size):
    rc.subgrid_to_facet_dft(dft, xA_size, yB_size, thin=3)


def test_subgrid_to_facet_dft_2d(dft):
    rc.subgrid_to_facet_dft_2d(dft)
