"""
Drop-in at the core-object boundary, checked against the REAL reference when it is mounted
(build container only; skipped on the GPU box): the reference's own, unmodified task bodies
(`api_helper.extract_column`, `sum_and_finish_subgrid`, `prepare_and_split_subgrid`,
`accumulate_column`, `accumulate_facet`, `finish_facet`) are driven once with the reference's
numpy `SwiftlyCore` and once with this repo's core object bound to the host-emulated kernels;
the results must agree.  This is the seat `SwiftlyConfig(backend=...)` fills
(reference api.py:137-143).
"""

import os
import sys

import numpy
import pytest

from tests import parity_cases as pc
from tests.emu_support import emu_core_class

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden  # pylint: disable=import-error,import-outside-toplevel

    api, api_helper, core_mod, _ = make_golden.import_reference()
    return api, api_helper, core_mod


def test_reference_task_bodies_run_on_our_core(ref):
    api, api_helper, core_mod = ref
    W, N, yB, yN, xA, xM = 13.5625, 256, 96, 128, 52, 64
    ref_core = core_mod.SwiftlyCore(W, N, xM, yN)
    our_core = emu_core_class()(W, N, xM, yN)
    facet_cfgs = api_helper.make_full_cover_config(N, yB, api.FacetConfig)
    sg_cfgs = api_helper.make_full_cover_config(N, xA, api.SubgridConfig)
    rng = numpy.random.default_rng(77)
    facets = [pc.rand_c(rng, yB, yB) for _ in facet_cfgs]

    def forward_backward(core):
        BF_F = [core.prepare_facet(f, fc.off0, axis=0) for f, fc in zip(facets, facet_cfgs)]
        sg = sg_cfgs[7]
        NMBF_BF = [api_helper.extract_column(core, bf, sg.off0, fc.off1)
                   for bf, fc in zip(BF_F, facet_cfgs)]
        contribs = [core.extract_from_facet(nb, sg.off1, axis=1) for nb in NMBF_BF]
        subgrid = api_helper.sum_and_finish_subgrid(core, contribs, facet_cfgs, sg)
        pieces = api_helper.prepare_and_split_subgrid(core, subgrid, [sg.off0, sg.off1], facet_cfgs)
        cols = [api_helper.accumulate_column(core, p, None, sg.off1) for p in pieces]
        accs = [api_helper.accumulate_facet(core, c, None, fc, sg.off0)
                for c, fc in zip(cols, facet_cfgs)]
        back = [api_helper.finish_facet(core, a, fc) for a, fc in zip(accs, facet_cfgs)]
        return subgrid, back

    sg_ref, back_ref = forward_backward(ref_core)
    sg_our, back_our = forward_backward(our_core)
    pc.close(sg_our, sg_ref, rtol=1e-12, what="subgrid through the reference's task bodies")
    for a, b in zip(back_our, back_ref):
        pc.close(a, b, rtol=1e-11, what="facet through the reference's task bodies")


def test_reference_unit_test_functions_accept_our_core(ref):
    """The reference's own tests/test_core.py test functions, run with our core class patched in
    for the 'numpy' backend (emulated kernels; thinned parameter set)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "ref_test_core", "/root/reference/tests/test_core.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cls = emu_core_class()
    mod.make_core = lambda pars, backend="numpy": cls(
        pars["W"], pars["N"], pars["xM_size"], pars["yN_size"])
    mod.test_base_params_fundamental("numpy")
    mod.test_base_params_derived("numpy")
    mod.test_base_params_check_params("numpy")
    mod.test_facet_to_subgrid_dft_2d("numpy")
    mod.test_subgrid_to_facet_dft_2d("numpy")


def test_unmodified_reference_native_backend_runs_on_our_library(ref):
    """The reference's `SwiftlyCoreFunc` (backend="ska_sdp_func") binds
    `ska_sdp_func.fourier_transforms.swiftly.Swiftly`; with our ska_sdp_func-shaped adapter in
    that seat (emulated kernels here) the reference's own unit tests pass UNMODIFIED, strided
    transposed views (axis 0) included."""
    import importlib.util

    from ska_sdp_distributed_fourier_transform_b200 import _lib, sdp_func_compat
    from tests.emu_support import emu_core_class

    emu_core_class()  # builds / loads the emulated library
    emu_cls = emu_core_class()
    lib = emu_cls(13.5625, 256, 64, 128)._lib

    class EmuSwiftly(sdp_func_compat.Swiftly):
        def __init__(self, N, yN_size, xM_size, W):
            real = _lib.load
            _lib.load = lambda path=None: lib
            try:
                super().__init__(N, yN_size, xM_size, W)
            finally:
                _lib.load = real

    native = sys.modules["ska_sdp_func.fourier_transforms.swiftly"]
    old = native.Swiftly
    native.Swiftly = EmuSwiftly
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_test_core_native", "/root/reference/tests/test_core.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        backend = "ska_sdp_func"
        mod.test_base_params_fundamental(backend)
        mod.test_base_params_derived(backend)
        mod.test_base_params_check_params(backend)
        mod.test_facet_to_subgrid_dft_2d(backend)
        mod.test_subgrid_to_facet_dft_2d(backend)
        mod.test_facet_to_subgrid_basic(227, 415, backend)
        mod.test_subgrid_to_facet_basic(228, 416, backend)
    finally:
        native.Swiftly = old
