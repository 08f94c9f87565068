"""The demo command-line drivers (scripts/) run on the GPU and meet the reference's bounds."""

import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_demo_api_round_trip():
    from ska_sdp_distributed_fourier_transform_b200 import SWIFT_CONFIGS

    demo = load("demo_api")
    errors = demo.demo_api(SWIFT_CONFIGS["1k[1]-n512-256"], 20, 1, 1, 1, True)
    assert max(errors) < 3e-10  # reference tests/test_api.py:125
    errors = demo.demo_api(SWIFT_CONFIGS["4k[1]-n2k-512"], 5, 2, 2, 10, False)
    assert max(errors) < 1e-7


def test_demo_sparse_facet_cover():
    from ska_sdp_distributed_fourier_transform_b200 import SWIFT_CONFIGS

    demo = load("demo_sparse_facet")
    offs = demo.disc_cover_offsets(4096, 704, 2.12 * 704)
    assert len(offs) == 7
    assert demo.demo(SWIFT_CONFIGS["4k[1]-n2k-512"], 2.12, 10, 20) < 1e-12
