"""
SwiftlyForward / SwiftlyBackward host logic and fused-kernel index algebra on the
host-emulated kernels (no GPU): compared with the reference-generated golden 2-D
fixture and with the oracle.  TEST TOOLING -- see tests/test_emu_parity.py.
"""

import pytest

from ska_sdp_distributed_fourier_transform_b200 import SwiftlyConfig
from tests import api_cases
from tests.emu_support import emu_core_class


def make_config(W, N, yB, yN, xA, xM, **kw):
    core = emu_core_class()(W, N, xM, yN, **kw)
    return SwiftlyConfig(W=W, fov=1.0, N=N, yB_size=yB, yN_size=yN, xA_size=xA, xM_size=xM,
                         core=core)


def test_emu_fused_ops_vs_oracle():
    api_cases.case_fused_ops_vs_oracle(make_config)


def test_emu_forward_backward_vs_reference_golden(golden_2d):
    api_cases.case_forward_backward_vs_reference_golden(make_config, golden_2d)


def test_emu_sparse_facets_shuffled_subgrids():
    api_cases.case_sparse_facets_shuffled_subgrids(make_config)


@pytest.mark.parametrize("lru_forward,lru_backward,shuffle", [(1, 1, False), (2, 2, True)])
def test_emu_api_round_trip(lru_forward, lru_backward, shuffle):
    api_cases.case_api_round_trip(make_config, lru_forward, lru_backward, shuffle)


def test_emu_fused_backward_ops_vs_oracle():
    api_cases.case_fused_backward_ops_vs_oracle(make_config)
    # yN through the 2 x yN/2 split kernel, and a non-power-of-two parameter set (split-F)
    api_cases.case_fused_backward_ops_vs_oracle(make_config, force_split=True)
    api_cases.case_fused_backward_ops_vs_oracle(make_config, W=11.0, N=1280, yB=440, yN=640,
                                                xA=280, xM=320)


def test_emu_many_sources():
    api_cases.case_many_sources(make_config)


@pytest.mark.parametrize("sg_variant", [0, 1, 2, 5, 16])
def test_emu_forward_backward_vs_oracle_kernel_variants(sg_variant):
    """Every variant of the fused subgrid kernel (0: two thread groups + TMA tensor stores,
    1: round-1 kernel, 2: two groups with the LSU token, 5: two groups, direct stores,
    16: first round with complex exchanges through the accumulator)."""
    api_cases.case_forward_backward_vs_oracle(make_config, sg_variant=sg_variant)


@pytest.mark.parametrize("force_split", [1, 2, 3, 4, 5, 6])
def test_emu_forward_split_k2_kernels(force_split):
    """K2 through the split kernels that serve yN = 16384 on the GPU, forced at yN = 512:
    1 = 2 x 256 (one thread group, E parked in the scratch), 2 = 4 x 128 with two thread groups
    on the same TMA-staged, swizzled row and a CTA-wide combine (the default at yN = 16384),
    3 = two fully independent groups (DIF across the groups, DIT within), 4 = the same
    decomposition with the intermediate results parked in (emulated) tensor memory and the
    groups swapping halves through it before 32-byte pair stores, 5 = tensor-memory parking
    with decimation in time across the groups as well (unit-stride output streams), 6 = the
    same with group 1's stores of a line deferred into the next line (kept half in tensor
    memory, bar.arrive / bar.sync hand-over of the parked half)."""
    import numpy

    from oracle.swiftly_oracle import OracleCore, forward_reference_order
    from ska_sdp_distributed_fourier_transform_b200 import FacetConfig, SubgridConfig, SwiftlyForward
    from tests import parity_cases as pc

    # (yB / yN = 0.5 like the BASELINE configs: well conditioned, see tests/test_gpu_parity.py)
    W, N, yB, yN, xA, xM = 13.5625, 1024, 256, 512, 128, 256
    cfg = make_config(W, N, yB, yN, xA, xM, force_split=force_split)
    oracle = OracleCore(W, N, xM, yN)
    rng = numpy.random.default_rng(force_split)
    offs = [(0, 256), (256, -256), (-256, 0)]
    facets = [pc.rand_c(rng, yB, yB), pc.rand_c(rng, yB, yB), pc.rand_c(rng, yB - 1, yB - 1)]
    sgs = [SubgridConfig(128, -128, xA), SubgridConfig(128, 384, xA), SubgridConfig(-384, 0, xA)]
    fwd = SwiftlyForward(cfg, [(FacetConfig(a, b, f.shape[0]), f) for (a, b), f in zip(offs, facets)])
    got = [fwd.get_subgrid_task(sg).result() for sg in sgs]
    # (the oracle's driver takes one facet size: run the odd-sized facet separately and add)
    ref = forward_reference_order(oracle, facets[:2], offs[:2], [(s.off0, s.off1) for s in sgs], xA)
    ref2 = forward_reference_order(oracle, facets[2:], offs[2:], [(s.off0, s.off1) for s in sgs], xA)
    for a, b, c in zip(got, ref, ref2):
        assert numpy.abs(a - (b + c)).max() <= 1e-12 * numpy.abs(b + c).max()
    # only the facets whose rows are whole 128-byte chunks: staged by swizzled tensor loads
    fwd = SwiftlyForward(cfg, [(FacetConfig(a, b, yB), f) for (a, b), f in zip(offs[:2], facets)])
    for sg, b in zip(sgs, ref):
        a = fwd.get_subgrid_task(sg).result()
        assert numpy.abs(a - b).max() <= 1e-12 * numpy.abs(b).max()
    # persistent CTAs that walk SEVERAL lines each (grid capped at 5 CTAs: 51-52 lines per CTA
    # -- staging refills, tensor-memory double buffering, stores deferred into the next line)
    import ctypes

    cfg.core._lib.swiftly_b200_debug_max_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int]
    cfg.core._lib.swiftly_b200_debug_max_blocks(cfg.core._plan, 5)
    fwd = SwiftlyForward(cfg, [(FacetConfig(a, b, yB), f) for (a, b), f in zip(offs[:2], facets)])
    for sg, b in zip(sgs, ref):
        a = fwd.get_subgrid_task(sg).result()
        assert numpy.abs(a - b).max() <= 1e-12 * numpy.abs(b).max()


@pytest.mark.parametrize("force_split", [1, 2, 3, 4, 5, 6])
def test_emu_k2_forms_long_and_odd_rows(force_split):
    """``extract_columns`` of every K2 form (see test_emu_forward_split_k2_kernels) against the
    oracle on rows LONGER than yN / 2 (both halves of the DIF forms' first step are non-zero) and
    of odd length (linear bulk copies instead of swizzled tensor loads), a few lines per CTA."""
    import ctypes

    import numpy
    import torch

    from oracle.swiftly_oracle import OracleCore
    from tests import parity_cases as pc

    W, N, yB, yN, xA, xM = 13.5625, 1024, 256, 512, 128, 256
    cfg = make_config(W, N, yB, yN, xA, xM, force_split=force_split)
    core = cfg.core
    core._lib.swiftly_b200_debug_max_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int]
    core._lib.swiftly_b200_debug_max_blocks(core._plan, 7)
    oracle = OracleCore(W, N, xM, yN)
    rng = numpy.random.default_rng(40 + force_split)
    sg_off0 = 3 * xA
    for sizes, offs in (((320, 320), (0, 256)), ((301, 257), (-256, 512))):
        bfs = [pc.rand_c(rng, yN, fs) for fs in sizes]
        refs = [oracle.prepare_facet(oracle.extract_from_facet(bf, sg_off0, axis=0), off1, axis=1)
                for bf, off1 in zip(bfs, offs)]
        outs = core.extract_columns([torch.from_numpy(bf.copy()) for bf in bfs], sg_off0, list(offs))
        for o, r in zip(outs, refs):
            assert numpy.abs(o.numpy() - r).max() <= 1e-12 * numpy.abs(r).max()
